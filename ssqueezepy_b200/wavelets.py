# -*- coding: utf-8 -*-
"""Frequency-domain wavelets (host side).

Mirrors the interface of the reference's `ssqueezepy/wavelets.py:14-470`
(`Wavelet`: `__call__`, `xifn`, `Psih`, `N`, `xi`, `dtype`, `fn`, `config`,
`name`) and the wavelet functions `morlet` (wavelets.py:498-527), `bump`,
`cmhat`, `hhhat` (533-607) and `gmw` (`_gmw.py:22-394`: L1 order 0 evaluated on the
device, L2 and higher orders on the host and uploaded as tables).

Role in this package: the *host* evaluations below feed the parameter logic
(scale bounds, centre frequencies, `ssq_freqs`) -- cheap, run once per call.
The per-scale sampling `psih(scale * xi)` of the hot path is done on the GPU
inside the fused kernels (Morlet / GMW-L1), or uploaded once as a table for
the other wavelets and custom functions (`device_spec()` tells which).
"""
import numpy as np
from types import FunctionType

from .configs import DEFAULTS

pi = np.pi
__all__ = ['Wavelet', 'morlet', 'gmw', 'bump', 'cmhat', 'hhhat',
           'center_frequency', 'xi_grid']


def xi_grid(N, scale=1., dtype=np.float64):
    """Radian frequencies of an N-point DFT, scaled: `[0..N//2, -(N-1)//2..-1] *
    2*pi/N * scale`, computed in float64 and stored as `dtype`
    (reference: wavelets.py:473-484)."""
    idx = np.arange(N, dtype=np.float64)
    idx[N // 2 + 1:] -= N
    return (idx * (scale * (2 * pi) / N)).astype(dtype)


def _as_dtype_scalars(dtype, *vals):
    t = np.dtype(dtype).type
    return [t(v) for v in vals]


# ---- wavelet function factories (return fn(w) working in `dtype`) ------------
def morlet(mu=None, dtype=None):
    mu = DEFAULTS['morlet']['mu'] if mu is None else mu
    dtype = DEFAULTS['morlet']['dtype'] if dtype is None else dtype
    cs = (1 + np.exp(-mu**2) - 2 * np.exp(-3 / 4 * mu**2)) ** (-.5)
    ks = np.exp(-.5 * mu**2)
    mu_t, ks_t, half, amp = _as_dtype_scalars(dtype, mu, ks, -.5,
                                              np.sqrt(2) * cs * pi**.25)

    def fn(w):
        w = np.atleast_1d(np.asarray(w, dtype=dtype))
        d = w - mu_t
        return amp * (np.exp(half * (d * d)) - ks_t * np.exp(half * (w * w)))
    fn.kind, fn.params = 'morlet', dict(mu=float(mu))
    return fn


def morsefreq(gamma, beta):
    """Peak radian frequency of a generalized Morse wavelet, beta > 0."""
    return np.exp((np.log(beta) - np.log(gamma)) / gamma)


def gmw(gamma=None, beta=None, norm=None, order=None, centered_scale=None,
        dtype=None):
    D = DEFAULTS['gmw']
    gamma = D['gamma'] if gamma is None else gamma
    beta = D['beta'] if beta is None else beta
    norm = D['norm'] if norm is None else norm
    order = D['order'] if order is None else order
    centered_scale = D['centered_scale'] if centered_scale is None else centered_scale
    dtype = D['dtype'] if dtype is None else dtype
    if gamma <= 0 or beta <= 0:
        raise ValueError("`gamma` and `beta` must be positive "
                         "(got %s, %s)" % (gamma, beta))
    if norm not in ('bandpass', 'energy'):
        raise ValueError("`norm` must be 'bandpass' or 'energy' (got %s)" % norm)
    if int(order) != order or order < 0:
        raise ValueError("`order` must be a non-negative integer (got %s)" % order)
    if norm != 'bandpass' or order != 0:
        return _gmw_general(gamma, beta, norm, int(order), centered_scale, dtype)
    wc = morsefreq(gamma, beta)
    g_t, b_t, wc_t, wcl_t = _as_dtype_scalars(dtype, gamma, beta, wc, np.log(wc))

    def fn(w):
        w = np.atleast_1d(np.asarray(w, dtype=dtype))
        if centered_scale:
            w = w * wc_t
        pos = (w >= 0)
        w = w * pos
        with np.errstate(divide='ignore', invalid='ignore'):
            out = 2 * np.exp(-b_t * wcl_t + wc_t**g_t + b_t * np.log(w) - w**g_t) * pos
        return out.astype(dtype)
    fn.kind = None if centered_scale else 'gmw'
    fn.params = dict(gamma=float(gamma), beta=float(beta))
    return fn


def _gmw_general(gamma, beta, norm, order, centered_scale, dtype):
    """Generalized Morse wavelets beyond the L1 order-0 case (Olhede & Walden; reference
    `_gmw.py:228-394`): energy (L2) normalisation and orders k >= 1, whose spectrum is
    the order-0 shape w^beta exp(-w^gamma) times a generalized Laguerre polynomial in
    2 w^gamma.  Evaluated on the host in `dtype`; the transform uploads it as a table."""
    from scipy.special import gammaln
    r = (2 * beta + 1) / gamma
    wc = morsefreq(gamma, beta)
    dt_ = np.dtype(dtype).type
    if order == 0:                                # L2: unit energy
        # log of sqrt(2 pi gamma 2^r / Gamma(r)); kept in log form: Gamma(r) and w^beta
        # overflow float32 for the default beta = 60
        amp = dt_(0.5 * (np.log(2. * pi * gamma) + r * np.log(2.) - gammaln(r)))
        coef = None
    else:
        c = r - 1
        m = np.arange(order + 1)
        # Laguerre coefficients (-1)^m C(k + c, k - m) / m!
        lag = ((-1.)**m * np.exp(gammaln(order + c + 1) - gammaln(c + m + 1)
                                 - gammaln(order - m + 1) - gammaln(m + 1)))
        if norm == 'bandpass':
            scale = 2 * np.sqrt(np.exp(gammaln(r) + gammaln(order + 1) - gammaln(order + r)))
        else:
            scale = np.sqrt(2 * pi * gamma * 2**r *
                            np.exp(gammaln(order + 1) - gammaln(order + r)))
        coef = (lag * scale).astype(dtype)
        amp = None
    g_t, b_t, wc_t = _as_dtype_scalars(dtype, gamma, beta, wc)

    def fn(w):
        w = np.atleast_1d(np.asarray(w, dtype=dtype))
        if centered_scale:
            w = w * wc_t
        pos = (w >= 0)
        w = w * pos
        with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
            if coef is None:
                out = np.exp(amp + b_t * np.log(w) - w**g_t) * pos
            else:
                poly = np.zeros(w.shape, dtype=dtype)
                for k_, ck in enumerate(coef):
                    poly += ck * (2 * w**g_t)**k_
                if norm == 'bandpass':
                    out = poly * np.exp(-b_t * np.log(wc_t) + wc_t**g_t
                                        + b_t * np.log(w) - w**g_t) * pos
                else:
                    out = poly * np.exp(b_t * np.log(w) - w**g_t) * pos
        return np.where(pos, out, 0).astype(dtype)
    fn.kind = None                                # table path
    fn.params = dict(gamma=float(gamma), beta=float(beta), norm=norm, order=order)
    return fn


def bump(mu=None, s=None, om=None, dtype=None):
    D = DEFAULTS['bump']
    mu = D['mu'] if mu is None else mu
    s = D['s'] if s is None else s
    om = D['om'] if om is None else om
    dtype = D['dtype'] if dtype is None else dtype
    cdt = np.dtype('complex64' if np.dtype(dtype) == np.float32 else 'complex128')
    c_om = cdt.type(2j * pi * om)
    norm = cdt.type(.443993816053287)
    edge = np.dtype(dtype).type(.999)

    def fn(w):
        w = np.atleast_1d(np.asarray(w, dtype=cdt))
        u = (w - cdt.type(mu)) / cdt.type(s)
        inside = np.abs(u) < edge
        with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
            out = (np.exp(c_om * w) / cdt.type(s) * inside
                   * np.exp(-1 / (1 - (u * inside)**2)) / norm)
        return out
    fn.kind, fn.params = None, dict(mu=mu, s=s, om=om)
    return fn


def cmhat(mu=None, s=None, dtype=None):
    D = DEFAULTS['cmhat']
    mu = D['mu'] if mu is None else mu
    s = D['s'] if s is None else s
    dtype = D['dtype'] if dtype is None else dtype
    mu_t, s_t, p, amp = _as_dtype_scalars(dtype, mu, s, 5 / 2,
                                          2 * np.sqrt(2 / 3) * pi**(-1 / 4))

    def fn(w):
        v = np.atleast_1d(np.asarray(w, dtype=dtype)) - mu_t
        return amp * (s_t**p * v**2 * np.exp(-s_t**2 * v**2 / 2) * (v >= 0))
    fn.kind, fn.params = None, dict(mu=mu, s=s)
    return fn


def hhhat(mu=None, dtype=None):
    D = DEFAULTS['hhhat']
    mu = D['mu'] if mu is None else mu
    dtype = D['dtype'] if dtype is None else dtype
    mu_t, half, amp = _as_dtype_scalars(dtype, mu, -1 / 2,
                                        2 / np.sqrt(5) * pi**(-1 / 4))

    def fn(w):
        v = np.atleast_1d(np.asarray(w, dtype=dtype)) - mu_t
        return amp * (v * (1 + v) * np.exp(half * v**2)) * (1 + np.sign(v))
    fn.kind, fn.params = None, dict(mu=mu)
    return fn


_FACTORIES = {'morlet': morlet, 'gmw': gmw, 'bump': bump, 'cmhat': cmhat,
              'hhhat': hhhat}
_NAMES = {'morlet': 'Morlet', 'gmw': 'GMW L1', 'bump': 'Bump', 'cmhat': 'Cmhat',
          'hhhat': 'Hhhat'}


class Wavelet:
    """Frequency-domain wavelet `psih(w)`; `Wavelet.SUPPORTED` lists built-ins.

        Wavelet('morlet'); Wavelet(('gmw', {'beta': 12, 'gamma': 3}));
        Wavelet(fn)  # custom `fn(w) -> psih` (evaluated on the host, uploaded
                     # once per call as a table)
    """
    SUPPORTED = set(_FACTORIES)
    DTYPES = {'float32', 'float64'}

    def __init__(self, wavelet='gmw', N=1024, dtype=None):
        self._dtype = None if dtype is None else self._dtype_str(dtype)
        self._Psih = self._Psih_N = self._Psih_scale = None
        self._set_fn(wavelet)
        self.N = N

    # -- construction ---------------------------------------------------------
    @staticmethod
    def _dtype_str(dtype):
        s = str(np.dtype(dtype)) if not isinstance(dtype, str) else dtype
        s = s.split('.')[-1]
        if s not in Wavelet.DTYPES:
            raise ValueError("`dtype` must be one of: float32, float64 (got %s)" % s)
        return s

    def _set_fn(self, wavelet):
        if isinstance(wavelet, FunctionType):
            self.fn, self.config, self._name = wavelet, {}, 'Custom'
            if self._dtype is None:
                out = np.asarray(wavelet(np.asarray([1.], dtype='float32')))
                self._dtype = ('float32' if out.dtype in (np.float32, np.complex64)
                               else 'float64')
            return
        msg = ("`wavelet` must be one of: (1) string name of supported wavelet; "
               "(2) tuple of (1) and dict of wavelet parameters (e.g. {'mu': 5}); "
               "(3) custom function taking `scale * xi` as input. (got: %s)"
               % str(wavelet))
        if isinstance(wavelet, tuple):
            if not (len(wavelet) == 2 and isinstance(wavelet[0], str)
                    and isinstance(wavelet[1], dict)):
                raise TypeError(msg)
            name, opts = wavelet[0].lower(), dict(wavelet[1])
        elif isinstance(wavelet, str):
            name, opts = wavelet.lower(), {}
        else:
            raise TypeError(msg)
        if name not in _FACTORIES:
            raise ValueError("`wavelet` must be one of: %s (got %s)"
                             % (', '.join(sorted(_FACTORIES)), name))
        if 'dtype' in opts:
            opts['dtype'] = self._dtype_str(opts['dtype'])
        if self._dtype is not None:
            opts['dtype'] = self._dtype
        full = dict(DEFAULTS[name])
        full.update(opts)
        self._dtype = full['dtype']
        self.fn = _FACTORIES[name](**full)
        self.config = full
        self._name = _NAMES[name]
        if name == 'gmw':                          # 'GMW L1' / 'GMW L2' [+ ' K' for order > 0]
            self._name = ('GMW L2' if full.get('norm') == 'energy' else 'GMW L1') + \
                (' K' if full.get('order', 0) else '')

    @classmethod
    def _init_if_not_isinstance(cls, wavelet, **kw):
        return wavelet if isinstance(wavelet, Wavelet) else cls(wavelet, **kw)

    # -- properties -------------------------------------------------------------
    @property
    def dtype(self):
        return self._dtype

    @property
    def name(self):
        return self._name

    @property
    def N(self):
        return self._N

    @N.setter
    def N(self, value):
        self._N = int(value)
        self._xi = xi_grid(self._N, 1., self.dtype)

    @property
    def xi(self):
        return self._xi

    def device_spec(self):
        """(kind, params) if the fused kernels evaluate this wavelet themselves,
        else None (a `psih[na, n_up]` table is uploaded)."""
        kind = getattr(self.fn, 'kind', None)
        if kind == 'morlet':
            return 'morlet', (self.fn.params['mu'],)
        if kind == 'gmw':
            return 'gmw', (self.fn.params['gamma'], self.fn.params['beta'])
        return None

    # -- evaluation -------------------------------------------------------------
    def xifn(self, scale=None, N=None):
        if scale is None:
            scale = 1.
        scale = np.asarray(scale, dtype=self.dtype)
        if scale.ndim > 1 and scale.squeeze().ndim > 1:
            raise ValueError("2D `scale` unsupported")
        if scale.ndim == 1 and scale.size > 1:
            scale = scale.reshape(-1, 1)
        base = self.xi if N is None else xi_grid(N, 1., self.dtype)
        return scale * base

    def __call__(self, w=None, *, scale=None, N=None, nohalf=True, imag_th=1e-8):
        if w is not None:
            psih = self.fn(np.asarray(w, dtype=self.dtype))
        else:
            psih = self.fn(self.xifn(scale, N))
        if not nohalf:
            psih = self._halve_nyquist(psih)
        if (np.iscomplexobj(psih) and imag_th is not None
                and psih.imag.sum() / psih.real.sum() < imag_th):
            psih = psih.real
        return psih

    @staticmethod
    def _halve_nyquist(psih):
        n = psih.shape[-1]
        if n % 2 == 0:
            psih[..., n // 2] /= 2
        return psih

    def Psih(self, scale=None, N=None, nohalf=True):
        """Cached `psih` at `scale`, `N` (host array; kept for API parity)."""
        n_given = N is not None
        N = N or self.N
        if scale is None and not n_given and self._Psih is not None:
            return self._Psih
        if (self._Psih is not None and N == self._Psih_N
                and len(scale) == len(self._Psih_scale)
                and np.allclose(scale, self._Psih_scale)):
            return self._Psih
        self._Psih = self(scale=scale, N=N, nohalf=nohalf)
        self._Psih_N = N
        self._Psih_scale = np.array(scale, copy=True)
        return self._Psih

    def support(self, rel_tol):
        """[w_lo, w_hi] outside which |psih(w)| < rel_tol * max|psih| (float64
        scan).  Used to skip frequency bins whose contribution is below the
        working precision.  Returns None when it cannot be bounded."""
        key = ('support', rel_tol)
        cache = self.__dict__.setdefault('_cache', {})
        if key in cache:
            return cache[key]
        res = None
        try:
            wav64 = (self if self.dtype == 'float64' or not self.config
                     else Wavelet((self._name_key(), {**self.config, 'dtype': 'float64'})))
            w_pk, pk = find_maximum(wav64.fn)
            W = max(12 * w_pk, 60.)
            grid = np.linspace(-W, W, 480001)
            vals = np.abs(np.asarray(wav64.fn(grid), dtype=np.complex128))
            vals[~np.isfinite(vals)] = 0
            nz = np.flatnonzero(vals > rel_tol * pk)
            if nz.size and nz[-1] < len(grid) - 2:
                step = grid[1] - grid[0]
                lo = grid[nz[0]] - 2 * step if nz[0] > 1 else -np.inf
                res = (lo, grid[nz[-1]] + 2 * step)
        except Exception:
            res = None
        cache[key] = res
        return res

    def time_support(self, rel_tol):
        """Two-sided time support per unit scale: |psi_a(t)| < rel_tol * max|psi_a|
        for |t| > c * a / 2 (measured on a float64 sampling of the wavelet at scale
        32 over 2^16 points).  None if it cannot be measured."""
        key = ('tsupport', rel_tol)
        cache = self.__dict__.setdefault('_cache', {})
        if key in cache:
            return cache[key]
        res = None
        try:
            wav64 = (self if self.dtype == 'float64' or not self.config
                     else Wavelet((self._name_key(), {**self.config, 'dtype': 'float64'})))
            a0, M = 32., 1 << 16
            psih = np.asarray(wav64(scale=a0, N=M), dtype=np.complex128).reshape(-1)
            mag = np.abs(np.fft.ifft(psih))
            idx = np.flatnonzero(mag > rel_tol * mag.max())
            tmax = int(np.minimum(idx, M - idx).max())
            if tmax < M // 4:
                res = 2. * (tmax + 1) / a0
        except Exception:
            res = None
        cache[key] = res
        return res

    def _name_key(self):
        return 'gmw' if self._name.startswith('GMW') else \
            {v: k for k, v in _NAMES.items()}[self._name]


# ---- searches used by the scale logic (reference algos.py:625-703) ---------
def find_maximum(fn, step_size=1e-3, steps_per_search=1e4, step_start=0,
                 step_limit=1000, min_value=-1):
    """Input value and value of the single maximum of |fn| (windowed scan)."""
    n = int(steps_per_search)
    width = int(n * step_size)
    best_val, best_arg, win = min_value, None, 0
    while True:
        lo = step_start + width * win
        xs = np.linspace(lo, lo + width, n, endpoint=False)
        ys = np.abs(np.asarray(fn(xs))).astype(np.float64)
        top = ys.max()
        if top > best_val:
            best_val, best_arg = top, xs[np.argmax(ys)]
        elif top < best_val:
            return best_arg, best_val
        win += 1
        if xs.max() > step_limit:
            raise ValueError("could not find function maximum with given "
                             "(step_size, steps_per_search, step_start, "
                             "step_limit, min_value)=({}, {}, {}, {}, {})".format(
                                 step_size, steps_per_search, step_start,
                                 step_limit, min_value))


def find_first_occurrence(fn, value, step_size=1e-3, steps_per_search=1e4,
                          step_start=0, step_limit=1000):
    """Earliest input at which |fn| attains `value` (windowed scan)."""
    n = int(steps_per_search)
    width = int(n * step_size)
    win, last = 0, False
    while True:
        lo = step_start + width * win
        xs = np.linspace(lo, lo + width, n, endpoint=False)
        if xs.max() > step_limit:
            last = True
            xs = np.clip(xs, None, step_limit)
        ys = np.abs(np.asarray(fn(xs))).astype(np.float64)
        gap = np.abs(ys - value)
        if np.any(gap <= np.abs(np.diff(ys)).max()):
            k = np.argmin(gap)
            return xs[k], ys[k]
        win += 1
        if last:
            raise ValueError("could not find input value to yield function "
                             "output value=%s" % value)


def _analytic_shift(xh):
    """Frequency axis ordered negatives first, then 0..Nyquist (even N keeps
    Nyquist on the right; reference wavelets.py:951-965)."""
    n = len(xh)
    if n % 2 == 0:
        return np.concatenate([xh[n // 2 + 1:], xh[:n // 2 + 1]])
    return np.fft.ifftshift(xh)


def center_frequency(wavelet, scale=None, N=1024, kind='energy', force_int=None,
                     viz=False):
    """Centre frequency (radians) of `wavelet`: 'energy', 'peak' or 'peak-ct'
    (reference wavelets.py:611-745).  `viz` is accepted and ignored."""
    from scipy import integrate
    if kind not in ('energy', 'peak', 'peak-ct'):
        raise ValueError("`kind` must be one of: energy, peak, peak-ct (got %s)" % kind)
    wavelet = Wavelet._init_if_not_isinstance(wavelet)
    if scale is None and kind != 'peak-ct':
        scale = (4 / pi) * find_maximum(wavelet.fn)[0]

    def sampled(sc):
        w = _analytic_shift(xi_grid(N, 1.))
        psih = np.asarray(wavelet(np.asarray(sc) * w))
        return w, np.abs(psih) ** 2

    if kind == 'peak-ct':
        return float(find_maximum(wavelet.fn)[0])
    if kind == 'peak':
        w, a2 = sampled(scale)
        return float(w[np.argmax(a2)])
    # 'energy'
    if force_int is False:
        sc0 = (4 / pi) * find_maximum(wavelet.fn)[0]
        w, a2 = sampled(sc0)
        return float(integrate.trapezoid(a2 * w) / integrate.trapezoid(a2)
                     * (sc0 / scale))
    w, a2 = sampled(scale)
    return float(integrate.trapezoid(a2 * w) / integrate.trapezoid(a2))
