"""Exploratory GPU parity probe (not a test): prints GPU-vs-oracle statistics."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
np.seterr(all="ignore")
import icnn_b200
from icnn_b200 import bundle_entropy as be
from oracle import synth, picnn_np, bundle_np

def r32(fg):
    def w(y):
        f, g = fg(y)
        return f.astype(np.float32).astype(np.float64), g.astype(np.float32).astype(np.float64)
    return w

def stats(tag, xg, xo, extra=""):
    d = np.abs(xg - xo).max(axis=1)
    print("%-28s max %.2e  median %.2e  frac>1e-4 %.3f  frac>1e-6 %.3f %s" % (tag, d.max(), np.median(d), np.mean(d > 1e-4), np.mean(d > 1e-6), extra), flush=True)

which = sys.argv[1:] or ["fg", "k2", "fused"]
if "fg" in which:
    for name, B in [("C1", 64), ("C3", 128), ("C4", 512), ("T", 128)]:
        cfg = synth.CONFIGS[name]
        p, x, y0 = synth.make_inputs(name, B=B)
        net = icnn_b200.PICNN.from_params(p)
        fgd = net.bind(x, affine=cfg["affine"])
        y = np.random.RandomState(0).uniform(0.05, 0.95, size=y0.shape)
        y = y.astype(np.float32).astype(np.float64)
        f, g = fgd(y)
        fo, go = picnn_np.make_fg(p, x, affine=cfg["affine"])(y)
        print("fg %-3s f relerr %.2e  g relerr %.2e (max|g| %.2f)" % (name, np.abs(f - fo).max() / np.abs(fo).max(), np.abs(g - go).max() / np.abs(go).max(), np.abs(go).max()), flush=True)

if "k2" in which:
    for name, B, nIter in [("C1", 64, 5), ("C3", 32, 10), ("T", 16, 10), ("C4", 256, 5), ("C2", 4, 30)]:
        cfg = synth.CONFIGS[name]
        p, x, y0 = synth.make_inputs(name, B=B)
        fg = r32(picnn_np.make_fg(p, x, affine=cfg["affine"]))
        variants = [("rl", "newton")] if cfg["variant"] == "rl" else [("lib", "pc"), ("dual", "newton"), ("lib", "newton")]
        for variant, solver in variants:
            if variant == "lib":
                o = bundle_np.solve_batch(fg, y0.copy(), nIter=nIter, variant="lib", solver="pc")
            else:
                o = bundle_np.solve_batch(fg, y0.copy(), nIter=nIter, variant=variant)
            t0 = time.time()
            r = be.solveBatch(fg, y0.copy(), nIter=nIter, solver=solver, variant=variant)
            dt = time.time() - t0
            cg = np.array([len(a) for a in r[1]]); co = np.array([len(a) for a in o[1]])
            stats("k2 %s %s/%s" % (name, variant, solver), r[0], o[0],
                  " counts== %.3f nIters== %.3f  (%.2fs)" % (np.mean(cg == co), np.mean(np.array(r[5]) == np.array(o[5])), dt))

if "fused" in which:
    for name, B, nIter in [("C1", 64, 5), ("C3", 64, 10), ("T", 32, 10), ("C4", 512, 5), ("C2", 8, 30)]:
        cfg = synth.CONFIGS[name]
        p, x, y0 = synth.make_inputs(name, B=B)
        fg64 = picnn_np.make_fg(p, x, affine=cfg["affine"])
        net = icnn_b200.PICNN.from_params(p)
        fgd = net.bind(x, affine=cfg["affine"])
        variant = cfg["variant"]
        o = bundle_np.solve_batch(fg64, y0.copy(), nIter=nIter, variant=variant)
        o32 = bundle_np.solve_batch(picnn_np.make_fg(p, x, affine=cfg["affine"], dtype=np.float32, out_dtype=np.float64), y0.copy(), nIter=nIter, variant=variant)
        r = be.solveBatch(fgd, y0.copy(), nIter=nIter, variant=variant)
        cg = np.array([len(a) for a in r[1]]); co = np.array([len(a) for a in o[1]])
        stats("fused %s %s" % (name, variant), r[0], o[0], " counts== %.3f nIters== %.3f" % (np.mean(cg == co), np.mean(np.array(r[5]) == np.array(o[5]))))
        stats("   oracle f32-fg vs f64-fg", o32[0], o[0])
        if variant == "lib":
            r = be.solveBatch(fgd, y0.copy(), nIter=nIter, variant=variant, solver="newton")
            stats("fused %s lib/newton" % name, r[0], o[0])
        yg, fgd_f = icnn_b200.gd.solve(fgd, y0, nIter=30, lr=0.01, momentum=0.3)
        yo, fo = picnn_np.momentum_gd(fg64, y0, 30, 0.01, 0.3)
        stats("gd %s" % name, yg, yo, " f relerr %.2e" % (np.abs(fgd_f - fo).max() / np.abs(fo).max()))
print("done")
