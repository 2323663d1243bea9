#!/usr/bin/env python
"""Measured parity of the CUDA path against the reference (golden fixtures) and the oracle (fp32 and
fp64) - the numbers behind the tolerances in tests/parity.py.  Run on the GPU box:
    python scripts/parity_report.py > gpurun_out/parity.json
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import svbrdf_estimation_b200 as S                    # noqa: E402
from oracle import reference_port as O                # noqa: E402
from tests import parity                               # noqa: E402
from tests.common import synthetic_maps               # noqa: E402


def ours(inp, tgt, cfg):
    x = torch.as_tensor(inp).cuda().requires_grad_(True)
    loss = S.rendering_loss_with_records(x, torch.as_tensor(tgt).cuda(), torch.as_tensor(cfg))
    loss.backward()
    renders = S.render_records(x.detach(), torch.as_tensor(cfg)).cpu().numpy()
    return float(loss.detach()), x.grad.cpu().numpy(), renders


AMBIGUOUS_DLOG = 1e-3     # |log difference| below which fp32 cannot be expected to get the SIGN of an L1 term right


def report(name, inp, tgt, cfg, l32, g32, r32, l64, g64, r64, r64_target=None):
    """r64_target (fp64 renders of the target maps) enables the sign-ambiguity analysis: the loss is an L1 of log
    renders, so d loss / d render = sign(l) / M; where the two renders agree to |l| < 1e-3 (but are not identical) an
    fp32 evaluation - the reference's own included - may pick the other sign, which flips that term's whole
    gradient.  One such term on a highlight pixel can carry the entire rel-L2 of a gradient group, so the groups are
    also reported over the pixels that have no such term."""
    loss, grad, renders = ours(inp, tgt, cfg)
    row = {"case": name, "shape": list(np.shape(inp)), "N": int(np.shape(cfg)[1]),
           "loss_rel_err_vs_ref64": abs(loss - float(l64)) / abs(float(l64)),
           "ref32_loss_rel_err_vs_ref64": abs(float(l32) - float(l64)) / abs(float(l64))}
    keep = None
    if r64_target is not None:
        l = np.abs(np.log(r64 + 0.1) - np.log(r64_target + 0.1))               # [B,N,3,H,W]
        amb = ((l > 0) & (l < AMBIGUOUS_DLOG)).any(axis=(1, 2))                # [B,H,W]
        keep = ~amb[:, None]
        row["sign_ambiguous_pixels"] = {"count": int(amb.sum()), "fraction": float(amb.mean()), "threshold_dlog": AMBIGUOUS_DLOG}
    for gname, s in parity.GROUPS:
        row["grad_" + gname] = {"ours_vs_ref32": parity.rel_l2(grad[:, s], g32[:, s]),
                                "ours_vs_ref64": parity.rel_l2(grad[:, s], g64[:, s]),
                                "ref32_vs_ref64": parity.rel_l2(g32[:, s], g64[:, s])}
        if keep is not None:
            row["grad_" + gname].update({"ours_vs_ref64_unambiguous": parity.rel_l2(grad[:, s] * keep, g64[:, s] * keep),
                                         "ref32_vs_ref64_unambiguous": parity.rel_l2(g32[:, s] * keep, g64[:, s] * keep)})
    if r64 is not None:
        rel = np.abs(renders - r64) / np.maximum(np.abs(r64), 1e-30)
        relr = np.abs(r32 - r64) / np.maximum(np.abs(r64), 1e-30)
        row["renders"] = {"ours_vs_ref32": parity.rel_l2(renders, r32), "ours_vs_ref64": parity.rel_l2(renders, r64),
                          "ref32_vs_ref64": parity.rel_l2(r32, r64),
                          "elem_rel_p99.9_ours": float(np.quantile(rel, 0.999)), "elem_rel_p99.9_ref32": float(np.quantile(relr, 0.999)),
                          "elem_rel_max_ours": float(rel.max()), "elem_rel_max_ref32": float(relr.max()),
                          "max_abs_dlog_ours": float(np.abs(np.log(renders.astype(np.float64) + 0.1) - np.log(r64 + 0.1)).max()),
                          "max_abs_dlog_ref32": float(np.abs(np.log(r32.astype(np.float64) + 0.1) - np.log(r64 + 0.1)).max())}
    return row


def main():
    rows = []
    gdir = os.path.join(ROOT, "tests", "golden")
    for fx in ("loss_bench", "loss_stress", "loss_n27", "loss_real"):
        g = dict(np.load(os.path.join(gdir, fx + ".npz")))
        rows.append(report("golden:" + fx, g["input"], g["target"], g["configs"], g["loss_f32"], g["grad_f32"],
                           g.get("renders_f32"), g["loss_f64"], g["grad_f64"], g.get("renders_f64")))
    for size, batch, stress in ((64, 3, False), (48, 2, True), (128, 4, False), (37, 2, False)):
        inp, tgt = synthetic_maps(batch, size, 11, stress), synthetic_maps(batch, size, 12, stress)
        torch.manual_seed(313)
        cfg = O.sample_loss_configs(batch)
        l64, g64 = O.rendering_loss_and_grad(inp.double(), tgt.double(), cfg)
        l32, g32 = O.rendering_loss_and_grad(inp, tgt, cfg)
        r64, r32 = O.render_batch(inp.double(), cfg).numpy(), O.render_batch(inp, cfg).numpy()
        rows.append(report("oracle:%dx%d B%d%s" % (size, size, batch, " stress" if stress else ""), inp.numpy(), tgt.numpy(),
                           cfg.numpy(), l32, g32.numpy(), r32, l64, g64.numpy(), r64, O.render_batch(tgt.double(), cfg).numpy()))
    if "--full" in sys.argv:
        # BASELINE.json shapes at full resolution: the oracle itself runs on the GPU here (same eager code, fp64 = ground
        # truth, fp32 = what the reference computes), in batch slices that fit its autograd graph
        for name, batch, size, nr, ns in (("C1 shape 256x256 B8 N9", 8, 256, 3, 6), ("C3 records 256x256 B4 N27", 4, 256, 9, 18),
                                          ("C4 resolution 1024x1024 B1 N9", 1, 1024, 3, 6)):
            inp, tgt = synthetic_maps(batch, size, 21), synthetic_maps(batch, size, 22)
            torch.manual_seed(313)
            cfg = O.sample_loss_configs(batch, nr, ns)
            dev = torch.device("cuda", 0)
            l64, g64 = O.rendering_loss_and_grad(inp.double().to(dev), tgt.double().to(dev), cfg)
            l32, g32 = O.rendering_loss_and_grad(inp.to(dev), tgt.to(dev), cfg)
            with torch.no_grad():
                r64, r32 = O.render_batch(inp.double().to(dev), cfg).cpu().numpy(), O.render_batch(inp.to(dev), cfg).cpu().numpy()
                r64t = O.render_batch(tgt.double().to(dev), cfg).cpu().numpy()
            rows.append(report("oracle-on-gpu:" + name, inp.numpy(), tgt.numpy(), cfg.numpy(), float(l32), g32.cpu().numpy(), r32,
                               float(l64), g64.cpu().numpy(), r64, r64t))
            # term-level treatment of the L1 sign (tests/parity.py flipped_term_analysis): which individual terms did the
            # kernels - and the reference's own fp32 run - take with the other sign, and what is left once only those are put back
            _, grad, _ = ours(inp.numpy(), tgt.numpy(), cfg.numpy())
            for who, g in (("ours", grad), ("ref32", g32.cpu().numpy())):
                res = parity.flipped_term_analysis(O, inp, tgt, cfg, g, device=dev)
                rows[-1]["flipped_terms_" + who] = {
                    "candidates_0<|l|<1e-3": res["candidates"], "terms": res["terms"], "flipped": res["flipped"],
                    "max_abs_l_of_a_flipped_term": res["max_abs_l_flipped"],
                    "rel_l2_vs_ref64_all_pixels_raw": {gn: parity.rel_l2(res["g_ours"][:, s], res["g64"][:, s]) for gn, s in parity.GROUPS},
                    "rel_l2_vs_ref64_all_pixels_flipped_terms_put_back": {gn: parity.rel_l2(res["g_corrected"][:, s], res["g64"][:, s]) for gn, s in parity.GROUPS}}
            torch.cuda.empty_cache()
    print(json.dumps(rows, indent=1))


if __name__ == "__main__":
    main()
