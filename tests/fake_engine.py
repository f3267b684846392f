"""TEST INFRASTRUCTURE: an oracle-backed stand-in for dp_gp_lvm_b200.engine.BoundEngine with the same packed
buffers and call sequence, on CPU tensors.  Only tests/test_distributed_cpu.py uses it, to run the model's
N-sharding / all-reduce orchestration under gloo where there is no GPU."""
import torch

from oracle import streaming as S


class OracleEngine:
    def __init__(self, n_local, d, q, m, b, mode, device=None, exp_variant=0, **kw):
        self.n, self.d, self.q, self.m, self.b, self.mode = n_local, d, q, m, b, mode
        self.ncols = d if mode == 0 else 1
        self.stats_len = b * m * m + b * m * self.ncols + d + 2
        self.launch_count = 0

    def _m(self):
        return "t" if self.mode == 0 else "d"

    def _pack(self, p2, p, yy, kl):
        if self.mode == 1:
            p = p[:, :, None]
        return torch.cat([p2.reshape(-1), p.reshape(-1), yy.reshape(-1), kl.reshape(-1)])

    def split_stats(self, stats):
        b, m, c, d = self.b, self.m, self.ncols, self.d
        o1 = b * m * m; o2 = o1 + b * m * c; o3 = o2 + d
        return stats[:o1].view(b, m, m), stats[o1:o2].view(b, m, c), stats[o2:o3], stats[o3:o3 + 2]

    def stats_fwd(self, mu, s, y, z, gamma, alpha, out=None):
        with torch.no_grad():
            p2, p = S.chunk_stats(z, mu, s, y, gamma, alpha.reshape(-1, 1), self._m())
            return self._pack(p2, p, (y ** 2).sum(0), torch.stack([(mu ** 2).sum(), (s - torch.log(s)).sum()]))

    @torch.enable_grad()
    def bound(self, n_total, stats, z, gamma, alpha, beta, wgt):
        st = stats.detach().clone().requires_grad_(True)
        leaves = [t.detach().clone().requires_grad_(True) for t in (z, gamma, alpha, beta)]
        z_, g_, a_, b_ = leaves
        w_ = None if wgt is None else wgt.detach().clone().requires_grad_(True)
        p2, p, yy, kl = self.split_stats(st)
        if self.mode == 1:
            p = p[:, :, 0]
        gp = S.bound_from_stats(n_total, self.d, p2, p, yy, kl[0], kl[1], z_, g_, a_.reshape(-1, 1), b_.reshape(-1, 1), w_, self._m())
        outs = torch.autograd.grad(gp, [st] + leaves + ([w_] if w_ is not None else []), allow_unused=True)
        outs = [torch.zeros_like(t) if o is None else o for o, t in zip(outs, [st] + leaves + ([w_] if w_ is not None else []))]
        return (gp.detach().reshape(1), outs[0], outs[1], outs[2], outs[3], outs[4], outs[5] if w_ is not None else None)

    @torch.enable_grad()
    def stats_bwd(self, mu, s, y, z, gamma, alpha, dstats):
        leaves = [t.detach().clone().requires_grad_(True) for t in (mu, s, z, gamma, alpha)]
        mu_, s_, z_, g_, a_ = leaves
        p2, p = S.chunk_stats(z_, mu_, s_, y, g_, a_.reshape(-1, 1), self._m())
        st = self._pack(p2, p, (y ** 2).sum(0), torch.stack([(mu_ ** 2).sum(), (s_ - torch.log(s_)).sum()]))
        return torch.autograd.grad((st * dstats).sum(), leaves)

    def check(self):
        pass

    def set_timing(self, e):
        pass


def scipy_polygamma(x, want_psi, want_tri):
    """CPU stand-in for dpgp_polygamma (dp_gp_lvm_b200/utils/special.py: POLYGAMMA_HOOK)."""
    from scipy.special import digamma, polygamma
    xn = x.detach().cpu().numpy()
    psi = torch.as_tensor(digamma(xn), dtype=x.dtype).reshape(x.shape) if want_psi else None
    tri = torch.as_tensor(polygamma(1, xn), dtype=x.dtype).reshape(x.shape) if want_tri else None
    return psi, tri
