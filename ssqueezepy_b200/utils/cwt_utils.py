# -*- coding: utf-8 -*-
"""CWT scale selection and validation (host side, run once per call).

Same behaviour as the reference's `ssqueezepy/utils/cwt_utils.py`:
`process_scales` (196-261), `infer_scaletype` (264-298), `make_scales` (301-372),
`cwt_scalebounds` (66-188), `logscale_transition_idx` (375-394),
`nv_from_scales` (397-409), `find_min_scale` / `find_max_scale` /
`find_max_scale_alt` / `find_downsampling_scale` (412-695),
`_process_fs_and_t` (698-718).  These functions decide the reassignment bin
edges, so tests compare their float64 outputs with the reference's by `==`.
"""
import numpy as np

from .common import WARN, p2up, pi, assert_is_one_of
from ..configs import DEFAULTS

__all__ = ['cwt_scalebounds', 'process_scales', 'infer_scaletype', 'make_scales',
           'logscale_transition_idx', 'nv_from_scales', 'find_min_scale',
           'find_max_scale', 'find_max_scale_alt', 'find_downsampling_scale',
           '_process_fs_and_t', 'adm_ssq', 'adm_cwt', 'integrate_analytic']


def _wav(wavelet):
    from ..wavelets import Wavelet
    return Wavelet._init_if_not_isinstance(wavelet)


def _to_numpy(a):
    if hasattr(a, 'detach'):
        return a.detach().cpu().numpy()
    return np.asarray(a)


# ---------------------------------------------------------------------------
def _process_fs_and_t(fs, t, N):
    """(dt, fs, t) from either a sampling rate or a uniform time vector."""
    if fs is not None and t is not None:
        WARN("`t` will override `fs` (both were passed)")
    if t is not None:
        if len(t) != N:
            raise Exception("`t` must be of same length as `x` "
                            "(%s != %s)" % (len(t), N))
        if not np.mean(np.abs(np.diff(t, 2, axis=0))) < 1e-7:
            raise Exception("Time vector `t` must be uniformly sampled.")
        fs = 1 / (t[1] - t[0])
    elif fs is None:
        fs = 1
    elif fs <= 0:
        raise ValueError("`fs` must be > 0")
    return 1 / fs, fs, t


def logscale_transition_idx(scales):
    """Index splitting a two-rate exponential array `[scales[:idx], scales[idx:]]`,
    or None if there is no single clean transition."""
    scales = _to_numpy(scales)
    curv = np.abs(np.diff(np.log(scales), 2, axis=0))
    idx = int(np.argmax(curv)) + 2
    peak = curv.max()
    curv[idx - 2] = 0
    th = 1e-14 if scales.dtype == np.float64 else 1e-6
    if not np.any(peak > 100 * np.abs(curv).mean()):
        return None
    if not np.all(np.abs(curv) < th):
        return None
    return idx


def nv_from_scales(scales):
    """Voices per octave at each scale of a `2**(k/nv)` array ([na, 1])."""
    scales = _to_numpy(scales)
    per_step = 1 / np.diff(np.log2(scales), axis=0)
    nv = np.vstack([per_step[:1], per_step])
    idx = logscale_transition_idx(scales)
    if idx is not None:
        jump = int(np.argmax(np.abs(np.diff(nv, axis=0)))) + 1
        assert jump == idx, "%s != %s" % (jump, idx)
    return nv


def infer_scaletype(scales):
    """('log' | 'linear' | 'log-piecewise', nv).  Thresholds depend on the array
    dtype, and callers pass the wavelet-dtype array on purpose."""
    scales = _to_numpy(scales)
    if not isinstance(scales, np.ndarray):
        raise TypeError("`scales` must be a numpy array (got %s)" % type(scales))
    if scales.dtype not in (np.float32, np.float64):
        raise TypeError("`scales.dtype` must be np.float32 or np.float64 "
                        "(got %s)" % scales.dtype)
    scales = scales.reshape(-1, 1)
    th_log = 4e-15 if scales.dtype == np.float64 else 8e-7
    th_lin = th_log * 1e3
    if np.mean(np.abs(np.diff(np.log(scales), 2, axis=0))) < th_log:
        nv = int(np.round(1 / np.diff(np.log2(scales), axis=0)[0].squeeze()))
        return 'log', nv
    if np.mean(np.abs(np.diff(scales, 2, axis=0))) < th_lin:
        return 'linear', None
    if logscale_transition_idx(scales) is None:
        raise ValueError("could not infer `scaletype` from `scales`; "
                         "`scales` array must be linear or exponential. "
                         "(got diff(scales)=%s..." % np.diff(scales, axis=0)[:4])
    return 'log-piecewise', nv_from_scales(scales)


# ---------------------------------------------------------------------------
def find_min_scale(wavelet, cutoff=1):
    """Scale at which the wavelet sampled at Nyquist equals |cutoff| * its peak
    (right of the peak for cutoff > 0, left for cutoff < 0)."""
    from ..wavelets import find_maximum, find_first_occurrence
    wavelet = _wav(wavelet)
    w_peak, peak = find_maximum(wavelet.fn)
    lo, hi = (w_peak, 10 * w_peak) if cutoff > 0 else (0, w_peak)
    w_cut, _ = find_first_occurrence(wavelet.fn, value=abs(cutoff) * peak,
                                     step_start=lo, step_limit=hi)
    return w_cut / pi


def find_max_scale(wavelet, N, bin_loc=1, bin_amp=1):
    """Scale at which the wavelet's amplitude is `bin_amp` of its maximum at DFT
    bin `bin_loc`."""
    from ..wavelets import center_frequency
    wavelet = _wav(wavelet)
    scale_c = (4 / pi) * center_frequency(wavelet, kind='peak-ct', N=N)
    psih = np.asarray(wavelet(scale=scale_c, N=N))[:N // 2 + 1]
    xi = np.asarray(wavelet.xifn(scale_c, N))
    top = int(np.argmax(psih))
    below = np.where(psih[:top] < psih.max() * bin_amp)[0]
    w_bin = xi[below[-1]]
    return scale_c * (w_bin / xi[bin_loc])


def find_max_scale_alt(wavelet, N, min_cutoff=.1, max_cutoff=.8):
    """Largest scale whose two lowest useful DFT bins straddle the wavelet's peak
    symmetrically (the 'minimal' preset)."""
    from ..wavelets import find_maximum, find_first_occurrence
    if max_cutoff <= 0 or min_cutoff <= 0:
        raise ValueError("`max_cutoff` and `min_cutoff` must be positive "
                         "(got %s, %s)" % (max_cutoff, min_cutoff))
    if max_cutoff <= min_cutoff:
        raise ValueError("must have `max_cutoff > min_cutoff` "
                         "(got %s, %s)" % (max_cutoff, min_cutoff))
    wavelet = _wav(wavelet)
    w_peak, peak = find_maximum(wavelet.fn)
    w_cut, _ = find_first_occurrence(wavelet.fn, value=min_cutoff * peak,
                                     step_start=0, step_limit=w_peak)
    left = np.arange(w_cut, w_peak, step=1 / N)
    spacing = (w_peak - left[:-1]) * 2
    steps = left[:-1] / spacing
    drops = np.where(np.diff(steps % 1) < -.8)[0]
    if len(drops) == 0:
        raise Exception("Failed to find suffciently-integer xi divisions; try "
                        "widening (min_cutoff, max_cutoff)")
    return spacing[drops[0] + 1] / (pi / (N / 2))


def find_downsampling_scale(wavelet, scales, span=5, tol=3, method='sum',
                            nonzero_th=.02, nonzero_tol=4., N=None, viz=False,
                            viz_last=False):
    """Index of the first scale past which consecutive wavelets are redundantly
    dense in frequency; None if never."""
    assert_is_one_of(method, 'method', ('any', 'all', 'sum'))
    N = N or 2048
    if isinstance(wavelet, np.ndarray):
        Psih = wavelet
    else:
        Psih = np.asarray(_wav(wavelet)(scale=scales, N=N))
    if len(Psih) != len(scales):
        raise ValueError("len(Psih) != len(scales) "
                         "(%s != %s)" % (len(Psih), len(scales)))
    Psih = Psih[:, :Psih.shape[1] // 2]
    n_groups = len(Psih) - span - 1
    i = 0
    for i in range(n_groups):
        grp = Psih[i:i + span]
        row_max = grp.max(axis=1)[:, None]
        if (grp > nonzero_th * row_max).sum() / span > nonzero_tol:
            continue
        peaks = np.where(grp == row_max)[1]
        joint = np.argmax(np.prod(grp, 0))
        dist = np.abs(peaks - joint)
        if method == 'any':
            dense = dist.max() < tol
        elif method == 'all':
            dense = not np.all(dist > tol)
        else:
            dense = dist.sum() < tol
        if dense:
            break
    return i if (i < n_groups - 1) else None


def cwt_scalebounds(wavelet, N, preset=None, min_cutoff=None, max_cutoff=None,
                    cutoff=None, bin_loc=None, bin_amp=None, use_padded_N=True,
                    viz=False):
    """(min_scale, max_scale) over which `wavelet` is well-behaved for length `N`.
    `preset` in ('maximal', 'minimal', 'naive', None)."""
    fallback = dict(min_cutoff=.6, max_cutoff=.8, cutoff=-.5)
    if preset is not None:
        if any((min_cutoff, max_cutoff, cutoff)):
            WARN("`preset` will override `min_cutoff, max_cutoff, cutoff`")
        elif preset == 'minimal' and any((bin_amp, bin_loc)):
            WARN("`preset='minimal'` ignores `bin_amp` & `bin_loc`")
        assert_is_one_of(preset, 'preset', ('maximal', 'minimal', 'naive'))
        if preset in ('naive', 'maximal'):
            min_cutoff = max_cutoff = None
            if preset == 'maximal':
                cutoff = -.5
        else:
            min_cutoff, max_cutoff, cutoff = (fallback['min_cutoff'],
                                              fallback['max_cutoff'],
                                              fallback['cutoff'])
    else:
        if min_cutoff is None:
            min_cutoff = fallback['min_cutoff']
        elif min_cutoff <= 0:
            raise ValueError("`min_cutoff` must be >0 (got %s)" % min_cutoff)
        if max_cutoff is None:
            max_cutoff = fallback['max_cutoff']
        elif max_cutoff < min_cutoff:
            raise ValueError("must have `max_cutoff > min_cutoff` "
                             "(got %s, %s)" % (max_cutoff, min_cutoff))
    bin_loc = bin_loc or (2 if preset == 'maximal' else None)
    bin_amp = bin_amp or (1 if preset == 'maximal' else None)
    cutoff = cutoff if cutoff is not None else fallback['cutoff']
    if preset == 'naive':
        return 1, N
    M = p2up(N)[0] if use_padded_N else N
    lo = find_min_scale(wavelet, cutoff=cutoff)
    if preset == 'maximal':
        hi = find_max_scale(wavelet, M, bin_loc=bin_loc, bin_amp=bin_amp)
    else:
        hi = find_max_scale_alt(wavelet, M, min_cutoff=min_cutoff,
                                max_cutoff=max_cutoff)
    return lo, hi


def make_scales(N, min_scale=None, max_scale=None, nv=32, scaletype='log',
                wavelet=None, downsample=None):
    """[na, 1] scales between the bounds: 'log' (`2**(k/nv)`), 'log-piecewise'
    (high scales decimated by `downsample`) or 'linear'."""
    if scaletype == 'log-piecewise' and wavelet is None:
        raise ValueError("must pass `wavelet` for `scaletype == 'log-piecewise'`")
    if min_scale is None and max_scale is None and wavelet is not None:
        min_scale, max_scale = cwt_scalebounds(wavelet, N, use_padded_N=True)
    else:
        min_scale = min_scale or 1
        max_scale = max_scale or N
    if downsample is None:
        downsample = DEFAULTS['make_scales']['downsample']
    downsample = int(downsample)
    na = int(np.ceil(nv * np.log2(max_scale / min_scale)))
    first = int(np.floor(nv * np.log2(min_scale)))
    last = first + na
    if scaletype in ('log', 'log-piecewise'):
        scales = 2 ** (np.arange(first, last) / nv)
        if scaletype == 'log-piecewise':
            cut = find_downsampling_scale(wavelet, scales)
            if cut is not None:
                scales = np.hstack([scales[:cut],
                                    scales[cut + downsample - 1::downsample]])
    elif scaletype == 'linear':
        lo, hi = 2 ** (first / nv), 2 ** (last / nv)
        scales = np.linspace(lo, hi, int(np.ceil(hi / lo)))
    else:
        raise ValueError("`scaletype` must be 'log' or 'linear'; "
                         "got: %s" % scaletype)
    return scales.reshape(-1, 1)


def process_scales(scales, N, wavelet=None, nv=None, get_params=False,
                   use_padded_N=True):
    """Validate an array of scales or build one from a string spec
    ('log', 'log-piecewise', 'linear', optionally ':maximal' / ':minimal').
    Returns `scales` ([na, 1]) or `(scales, scaletype, na, nv)`."""
    preset = None
    if isinstance(scales, str):
        if ':' in scales:
            scales, preset = scales.split(':')
        elif scales == 'log-piecewise':
            preset = 'maximal'
        assert_is_one_of(scales, 'scales', ('log', 'log-piecewise', 'linear'))
        if nv is None:
            nv = 32
        if wavelet is None:
            raise ValueError("must set `wavelet` if `scales` isn't array")
        scaletype = scales
    elif isinstance(scales, np.ndarray) or hasattr(scales, 'detach'):
        scales = _to_numpy(scales)
        if scales.squeeze().ndim != 1:
            raise ValueError("`scales`, if array, must be 1D "
                             "(got shape %s)" % str(scales.shape))
        scaletype, found_nv = infer_scaletype(scales)
        if scaletype == 'log':
            if nv is not None and found_nv != nv:
                raise Exception("`nv` used in `scales` differs from "
                                "`nv` passed (%s != %s)" % (found_nv, nv))
            nv = found_nv
        elif scaletype == 'log-piecewise':
            nv = found_nv
        scales = scales.reshape(-1, 1)
    else:
        raise TypeError("`scales` must be a string or Numpy array "
                        "(got %s)" % type(scales))
    if nv is not None and not isinstance(nv, np.ndarray):
        if not (nv > 0 and float(nv).is_integer()):
            raise ValueError("'nv' must be a positive integer (got %s)" % nv)
        nv = int(nv)
    if isinstance(scales, np.ndarray):
        return (scales, scaletype, len(scales), nv) if get_params else scales
    lo, hi = cwt_scalebounds(wavelet, N=N, preset=preset,
                             use_padded_N=use_padded_N)
    scales = make_scales(N, lo, hi, nv=nv, scaletype=scaletype, wavelet=wavelet)
    return (scales, scaletype, len(scales), nv) if get_params else scales


# ---- admissibility constants (inverse transforms) ------------------------------
def _first_below(a, th):
    """Index of the first entry below `th` (last index when there is none)."""
    hit = np.flatnonzero(a < th)
    return int(hit[0]) if len(hit) else len(a) - 1


def integrate_analytic(int_fn, nowarn=False):
    """Trapezoid integral over (0, inf) of a function that vanishes for negative
    arguments, has one peak and decays to the right -- the numerical scheme of the
    reference (`utils/cwt_utils.py:583-627`): a log-spaced piece on [1e-15, 0.1] plus
    a linear grid on [0.1, mx) whose right end is pushed out (1, 20, 80, 160) until the
    integrand has decayed below 1e-15 well inside it."""
    from scipy import integrate
    t0 = np.logspace(-15, -1, 1000)
    near_zero = integrate.trapezoid(int_fn(t0), t0)

    chosen = None
    for m, mx in zip((1, 1, 4, 8), (1, 20, 80, 160)):
        n = 10000 * m
        t = np.linspace(mx, .1, n, endpoint=False)[::-1].copy()
        arr = int_fn(t)
        peak = int(np.argmax(arr))
        cut = _first_below(np.abs(arr[peak:]), 1e-15) + peak
        chosen = (arr, t, cut)
        if (len(t) - cut > 1000 * m) and np.sum(np.abs(arr)) > 1e-5:
            break
    else:
        if near_zero < 1e-5:
            raise Exception("Could not find converging or non-negligibly"
                            "-valued bounds of integration for `int_fn`")
        elif not nowarn:
            WARN("Integrated only from 1e-15 to 0.1 in logspace")
    arr, t, cut = chosen
    return integrate.trapezoid(arr[:cut], t[:cut]) + near_zero


def _real_if_real(c):
    return c.real if abs(c.imag) < 1e-15 else c


_ADM_CACHE = {}


def _adm_cached(kind, wavelet, integrand):
    """The constants depend on the wavelet only: memoised for built-in wavelets (the
    integration samples the wavelet ~10^5 times, ~0.5 ms)."""
    wav = _wav(wavelet)
    key = None
    if wav.config:
        key = (kind, wav.name, wav.dtype, tuple(sorted((k, str(v)) for k, v in wav.config.items())))
        if key in _ADM_CACHE:
            return _ADM_CACHE[key]
    val = _real_if_real(integrate_analytic(lambda w: integrand(_to_numpy(wav.fn(w)), w)))
    if key is not None:
        if len(_ADM_CACHE) > 64:
            _ADM_CACHE.clear()
        _ADM_CACHE[key] = val
    return val


def adm_ssq(wavelet):
    """Synchrosqueezing admissibility constant: integral of conj(psih(w)) / w over
    w > 0 (reference `utils/cwt_utils.py:28-47`)."""
    return _adm_cached('ssq', wavelet, lambda p, w: np.conj(p) / w)


def adm_cwt(wavelet):
    """CWT admissibility constant: integral of |psih(w)|^2 / w over w > 0
    (reference `utils/cwt_utils.py:50-63`)."""
    return _adm_cached('cwt', wavelet, lambda p, w: np.conj(p) * p / w)
