#!/usr/bin/env python
"""Is bench.py's `cpu_baseline.reference_cost` (the numpy port run with the reference's dense np.diag matrices and
per-iteration prints, one thread) representative of the UNMODIFIED reference?  Times lib/bundle_entropy.solveBatch imported
from /root/reference (build container only) and the port in that mode on the same rows, one BLAS thread each, float32-
arithmetic fg.  Writes profiles/r02c_reference_cost_check.json.  CPU only, test infrastructure."""
import contextlib
import io
import json
import os
import sys
import time

import numpy as np
from threadpoolctl import threadpool_limits

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import bundle_np, picnn_np, synth  # noqa: E402
from oracle.gen_golden import REF, _load  # noqa: E402

CASES = [("C3", 16, 10), ("T", 8, 10), ("C2", 2, 6), ("C5", 1, 3)]      # (config, rows, iterations)


def main():
    ref = _load("ref_pc", os.path.join(REF, "lib/bundle_entropy.py"))
    out = {}
    for name, rows, its in CASES:
        p, x, y0 = synth.make_inputs(name, B=rows)
        fg = picnn_np.make_fg(p, x, dtype=np.float32, out_dtype=np.float64)
        rec = {"rows": rows, "iterations": its}
        with threadpool_limits(limits=1), np.errstate(all="ignore"), contextlib.redirect_stdout(io.StringIO()):
            for tag, fn in (("reference_unmodified", lambda: ref.solveBatch(fg, y0.copy(), nIter=its)),
                            ("port_reference_cost_mode", lambda: bundle_np.solve_batch(fg, y0.copy(), nIter=its, dense_diag=True, verbose=True)),
                            ("port_default", lambda: bundle_np.solve_batch(fg, y0.copy(), nIter=its))):
                t0 = time.perf_counter()
                r = fn()
                dt = time.perf_counter() - t0
                rec[tag] = {"seconds": round(dt, 3), "solves_per_s": round(rows * its / dt, 3)}
                rec[tag]["y_checksum"] = float(np.abs(r[0]).sum())
        rec["port_cost_mode_over_reference"] = round(rec["port_reference_cost_mode"]["seconds"] / rec["reference_unmodified"]["seconds"], 3)
        out[name] = rec
        sys.stderr.write("%s %s\n" % (name, json.dumps(rec)))
    with open(os.path.join(ROOT, "profiles", "r02c_reference_cost_check.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
