# -*- coding: utf-8 -*-
"""Default parameter values (the values the reference ships in
`ssqueezepy/configs.ini:7-40`).  One plain dict; there is no ini re-parsing per
call, no SSQ_GPU / SSQ_PARALLEL environment switch: this package always runs the
CUDA path."""

DEFAULTS = {
    'morlet': dict(mu=13.4, dtype='float32'),
    'bump':   dict(mu=5, s=1, om=0, dtype='float32'),
    'cmhat':  dict(mu=1, s=1, dtype='float32'),
    'hhhat':  dict(mu=5, dtype='float32'),
    'gmw':    dict(gamma=3, beta=60, norm='bandpass', order=0,
                   centered_scale=False, dtype='float32'),
    'stft':   dict(dtype='float32'),
    'make_scales': dict(downsample=4),
}


def USE_GPU():
    """Always True (kept for API familiarity; reference configs.py:142-147)."""
    return True


def IS_PARALLEL():
    """CPU-thread parallelism does not apply (reference configs.py:127-139)."""
    return False
