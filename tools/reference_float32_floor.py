#!/usr/bin/env python
"""How far the UNMODIFIED reference moves under float32 rounding of its own callback: lib/bundle_entropy.solveBatch
(imported from /root/reference, build container only) fed the float64 PICNN fg versus the same fg evaluated in float32
arithmetic -- (a) handed over in float64 arrays, (b) handed over as float32 arrays like the TensorFlow fetch of the
reference scripts (multi-label-cls/icnn_ebundle.py:218-221), where np.linalg.matrix_rank additionally scales its
tolerance with the row dtype.  This is the floor any float32 implementation of fg (TensorFlow's included) sits on;
the device path's distance to the float64 oracle (profiles/r02_parity.json) is to be read against it.
Writes profiles/r02c_reference_float32_floor.json.  CPU only, test infrastructure (uses oracle/)."""
import contextlib
import io
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import picnn_np, synth  # noqa: E402
from oracle.gen_golden import REF, _load  # noqa: E402

CASES = [("C3", 64), ("T", 64), ("C2", 16), ("C5", 4), ("C4", 256)]      # (config, rows) at the configs' own horizons


def stats(a, b):
    d = np.abs(a - b).max(axis=1)
    return {"max": float(d.max()), "median": float(np.median(d)), "frac_gt_1e-4": float((d > 1e-4).mean())}


def main():
    out = {"_what": __doc__.split("Writes")[0].strip().replace("\n", " ")}
    for name, rows in CASES:
        cfg = synth.CONFIGS[name]
        path = "RL/src/bundle_entropy.py" if cfg["variant"] == "rl" else "lib/bundle_entropy.py"
        ref = _load("ref_" + name, os.path.join(REF, path))
        p, x, y0 = synth.make_inputs(name, B=rows)
        res = {}
        t0 = time.time()
        for tag, kw in (("float64", {}), ("float32_arith_float64_rows", dict(dtype=np.float32, out_dtype=np.float64)),
                        ("float32_fetch", dict(dtype=np.float32, out_dtype=np.float32))):
            fg = picnn_np.make_fg(p, x, affine=cfg["affine"], **kw)
            with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
                r = ref.solveBatch(fg, y0.copy(), nIter=cfg["nIter"])
            res[tag] = (np.array(r[0], dtype=np.float64), [len(a) for a in r[1]], list(r[5]))
        y64 = res["float64"][0]
        rec = {"rows": rows, "nIter": cfg["nIter"], "n_y": cfg["n"], "module": path, "seconds": round(time.time() - t0, 1),
               "mean_active_rows_float64": float(np.mean(res["float64"][1]))}
        for tag in ("float32_arith_float64_rows", "float32_fetch"):
            rec[tag + "_vs_float64"] = dict(stats(res[tag][0], y64), mean_active_rows=float(np.mean(res[tag][1])),
                                            mean_nIters=float(np.mean(res[tag][2])))
        out[name] = rec
        print(name, json.dumps(rec))
    with open(os.path.join(ROOT, "profiles", "r02c_reference_float32_floor.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
