"""Developer aid: every configuration of the reference's test suite at FULL size (tests/inputs/full/, written by
scripts/make_test_inputs.py --full: the reference's scans, maps and iteration limits) through the twin main program,
timed end to end on one GPU: wall time of the whole run (table generation, fits with --fit, upload, solves, file
output), the number of D(omega,k) evaluations and set_k calls it needed.
    python scripts/full_configs.py [--out gpurun_out/full_configs.jsonl] [names...]"""
import io, json, os, sys, time, contextlib, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from alps_b200 import run

FULL = os.path.join(ROOT, "tests", "inputs", "full")
repeat = 1
if "--repeat" in sys.argv:
    repeat = int(sys.argv[sys.argv.index("--repeat") + 1])
    sys.argv.pop(sys.argv.index("--repeat") + 1)
names = [a for a in sys.argv[1:] if not a.startswith("--")]
out_path = "gpurun_out/full_configs.jsonl"
if "--out" in sys.argv:
    out_path = sys.argv[sys.argv.index("--out") + 1]
    names = [n for n in names if n != out_path]
if not names:
    names = sorted(f[:-3] for f in os.listdir(FULL) if f.endswith(".in") and not f.endswith("_dist.in"))
os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
# warm-up: CUDA context, module load
with contextlib.redirect_stdout(io.StringIO()):
    run.main([os.path.join(ROOT, "tests", "inputs", "test_map_small.in"), "--dist",
              os.path.join(ROOT, "tests", "inputs", "test_kpar_fast_dist.in"), "--out", tempfile.mkdtemp(), "--nproc", "4"])
with open(out_path, "w") as fh:
    for name in names:
        inp, dist = os.path.join(FULL, name + ".in"), os.path.join(FULL, name + "_dist.in")
        out = tempfile.mkdtemp()
        args = [inp, "--out", out, "--nproc", "4", "--fit"] + (["--dist", dist] if os.path.exists(dist) else [])
        dt = None
        for _ in range(repeat):      # --repeat N: best of N (the runs are deterministic, the box is not)
            buf = io.StringIO()
            t0 = time.perf_counter()
            try:
                with contextlib.redirect_stdout(buf):
                    rc = run.main(args)
            except Exception as e:      # report and go on
                rc = repr(e)
            t1 = time.perf_counter() - t0
            dt = t1 if dt is None else min(dt, t1)
        d_evals, set_k, memo, pref = getattr(run.main, "last_stats", (0, 0, 0, 0))
        files = sorted(os.listdir(out))
        rec = {"config": name, "rc": rc, "wall_s": dt, "D_evals": d_evals, "set_k_calls": set_k, "memo_hits": memo, "prefetched": pref,
               "disp_calls_per_s_end_to_end": (d_evals + memo - pref) / dt if dt > 0 else None, "files": len(files)}
        print(json.dumps(rec), flush=True)
        fh.write(json.dumps(rec) + "\n")
