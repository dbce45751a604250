"""Generates tests/golden/aggregation_cpu.npz by running the REFERENCE's own pure-PyTorch
Aggregation.py (imported from /root/reference, CPU) on seeded inputs: blend weights, their
autograd gradients, merge_final and expend_sigma.  Run in the build container only:
    python tools/make_golden_cpu.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _ref_import  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "aggregation_cpu.npz")


def make_hits(g, R, K, p_empty=0.3):
    """Sorted hit lists with sentinel-padded tails, like the fine kernel emits."""
    nvalid = torch.randint(0, K + 1, (R,), generator=g)
    nvalid[0] = 0
    nvalid[1] = K
    ln = torch.sort(torch.rand(R, K, generator=g) * 4 + 3, dim=1).values
    # some tight clusters so that erf is not saturated
    ln[:, 1::2] = ln[:, 0::2][:, : ln[:, 1::2].shape[1]] + torch.rand(R, ln[:, 1::2].shape[1], generator=g) * 0.05
    ln = torch.sort(ln, dim=1).values
    act = torch.rand(R, K, generator=g) * 4.6
    dsd = torch.exp(torch.rand(R, K, generator=g) * 6 + 1)
    idx = torch.randint(0, 500, (R, K), generator=g).int()
    k = torch.arange(K)[None]
    empty = k >= nvalid[:, None]
    ln[empty], act[empty], dsd[empty], idx[empty] = 1e10, 1e10, 0.0, -1
    return idx, act, ln, dsd


def main():
    ref = _ref_import.load()
    A = ref["Aggregation"]
    g = torch.Generator().manual_seed(1234)
    out = {}
    for tag, (R, K, occ) in {"k5": (64, 5, 1.0), "k20": (256, 20, 1.0), "k40_occ": (64, 40, 2.5)}.items():
        idx, act, ln, dsd = make_hits(g, R, K)
        act.requires_grad_(True); ln.requires_grad_(True); dsd.requires_grad_(True)
        w, idx_o, valid, ln_o = A.aggregation(idx.view(1, R, 1, K), act.view(1, R, 1, K), ln.view(1, R, 1, K),
                                              dsd.view(1, R, 1, K), occ)
        gw = torch.rand(w.shape, generator=g)
        (w * gw).sum().backward()
        attr = torch.rand(500, 3, generator=g)
        merged = A.merge_final(attr, w.detach(), idx.view(1, R, 1, K).clone(), valid)
        out.update({tag + "_idx": idx.numpy(), tag + "_act": act.detach().numpy(), tag + "_len": ln.detach().numpy(),
                    tag + "_dsd": dsd.detach().numpy(), tag + "_occ": np.float32(occ),
                    tag + "_weight": w.detach().view(R, K).numpy(), tag + "_valid": valid.view(R).numpy(),
                    tag + "_gw": gw.view(R, K).numpy(), tag + "_g_act": act.grad.numpy(),
                    tag + "_g_len": ln.grad.numpy(), tag + "_g_dsd": dsd.grad.numpy(),
                    tag + "_attr": attr.numpy(), tag + "_merged": merged.view(R, 3).numpy()})
    # the sentinel example quoted in SURVEY.md 7(4)
    s = torch.tensor([0.5, 1.0, 2.0])
    out["expend_1"] = A.expend_sigma(s).numpy()
    out["expend_2"] = A.expend_sigma(torch.tensor([[1., 2., 3.], [4., 5., 6.]])).numpy()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items() if k.startswith("k5")})


if __name__ == "__main__":
    main()
