"""Outer AO-ADMM loop of the reference CLI (scripts/factorize.py:176-318) as a reusable object.

One `LayerSolver` owns one layer's problem on one GPU: the three unfoldings of W (made once),
the factors, the scaled duals and the loss histories.  A sweep enqueues, per mode, the
Gram-Hadamard, the MTTKRP, the persistent ADMM kernel and the re-projection, then two
reconstruction-error reductions; the host synchronises once per sweep to apply the reference's
stop rules (:259-263 / :303-307)."""
import numpy as np
import torch

from . import _native


class LayerSolver:
    def __init__(self, weight, factors, bits, qscheme, max_iter_admm=1000, eps=1e-8, tol=1e-5,
                 num_attempts=200, mttkrp_precision=0, init_is_random=True):
        _native.require_cuda(weight)
        assert weight.ndim in (2, 3), "Incorrect number of dimentions in weight tensor"
        self.W = _native.f32c(weight)
        self.N = self.W.ndim
        self.bits, self.qscheme = int(bits), qscheme
        _native.qscheme_id(qscheme)
        self.max_iter_admm, self.eps, self.tol = int(max_iter_admm), float(eps), float(tol)
        self.num_attempts, self.mttkrp_precision = int(num_attempts), int(mttkrp_precision)
        dev = self.W.device
        self.factors = [_native.f32c(f).to(dev).clone() for f in factors]
        self.duals = [torch.zeros_like(f) for f in self.factors]       # :209-212 / :272-273
        self.factors_q = [None] * self.N
        self.loss_hist, self.loss_quant_hist = [], []
        self.reports = []
        if self.N == 3:
            I, J, K = self.W.shape
            self.unfoldings = [self.W.reshape(I, J * K), _native.unfold3(self.W, 1), _native.unfold3(self.W, 2)]
        else:
            self.unfoldings = [self.W, self.W.t().contiguous()]
        if not init_is_random:                                          # :192-201
            fq = [_native.project(f, self.bits, qscheme, self.num_attempts)[0] for f in self.factors]
            self.loss_hist.append(self._error(self.factors))
            self.loss_quant_hist.append(self._error(fq))

    # ---- pieces
    def _others(self, mode):
        return [self.factors[k] for k in range(self.N) if k != mode]

    def _error_sums(self, fac):
        if self.N == 3:
            return _native.recon_error_sums(self.unfoldings[0], fac[0], fac[1], fac[2])
        return _native.recon_error_sums(self.unfoldings[0], fac[0], fac[1], None)

    @staticmethod
    def _finish_error(sums):
        num, den = (np.float32(v) for v in sums.cpu().numpy())
        return float(np.sqrt(np.float32(num / den)))                    # source/admm.py:15 in float32

    def _error(self, fac):
        return self._finish_error(self._error_sums(fac))

    def update_mode(self, mode, codes=None):
        """One ALS step for `mode`: scripts/factorize.py:215-224 (and the two analogous blocks)."""
        others = self._others(mode)
        G = _native.gram_hadamard(others[0], others[1] if self.N == 3 else None)
        F = _native.mttkrp(self.unfoldings[mode], others[0], others[1] if self.N == 3 else None,
                           self.mttkrp_precision)
        rep = _native.admm_iteration_inplace(self.factors[mode], self.duals[mode], F, G, self.max_iter_admm,
                                             self.eps, self.bits, self.qscheme, self.num_attempts, codes)
        self.reports.append(rep)
        self.factors_q[mode] = _native.project(self.factors[mode], self.bits, self.qscheme, self.num_attempts)[0]
        return F, G

    def sweep(self):
        """One outer iteration (:214-258); returns (error, quantized_error) after one host sync."""
        self.reports = []
        for mode in range(self.N):
            self.update_mode(mode)
        s1 = self._error_sums(self.factors)                             # :246-248
        s2 = self._error_sums(self.factors_q)                           # :249-253
        for rep in self.reports:
            _native.read_report(rep)                                    # LinAlgError on a non-PD system
        err, errq = self._finish_error(s1), self._finish_error(s2)
        self.loss_hist.append(err)
        self.loss_quant_hist.append(errq)
        return err, errq

    def should_stop(self):
        """Stop rules of scripts/factorize.py:259-263 (3-D) and :303-307 (2-D)."""
        h = self.loss_hist
        if len(h) > 1 and abs(h[-2] - h[-1]) < self.tol:
            return True
        back = 5 if self.N == 3 else 10
        if len(h) > 10 and h[-1] - h[-back] > 1e-3:
            return True
        return False

    def run(self, max_iter_als, progress=None):
        sweeps = 0
        it = range(max_iter_als) if progress is None else progress(range(max_iter_als))
        for _ in it:
            self.sweep()
            sweeps += 1
            if self.should_stop():
                break
        return sweeps

    def inner_iterations_per_sweep(self):
        return self.N * max(self.max_iter_admm - 1, 0)


def rank_from_reduction_rate(weight, reduction_rate):
    """scripts/factorize.py:157-158."""
    return int(weight.numel() / sum(list(weight.shape)) / reduction_rate)


def layer_weight_as_tensor(weight):
    """Intended reshape of scripts/factorize.py:138-145 / scripts/calibrate.py:178-184:
    conv (Cout,Cin,kh,kw) -> (Cout,Cin,kh*kw); 1x1 conv -> (Cout,Cin)."""
    if weight.ndim == 4:
        if tuple(weight.shape[2:]) == (1, 1):
            return weight.reshape(weight.shape[0], weight.shape[1])
        return weight.reshape(weight.shape[0], weight.shape[1], -1)
    return weight
