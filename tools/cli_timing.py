"""Wall-clock of the drop-in CLI on a configs[1]-shaped database written to disk (.skm/.skd):
`sketchlib dist db --knn 50` and dense `sketchlib dist db -o out.txt --threads T` (text tail included)."""
import os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import oracle
from helpers import write_skm_skd
from sketchlib.rust_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
kmers = list(range(15, 31, 2))
bins = synth.as_u64_numpy(synth.make_bins(n, 16, kmers, seed=1, device="cuda"))
db = oracle.Db(bins, n, kmers, 16, names=[f"genome_{i:06d}.fa" for i in range(n)])
d = tempfile.mkdtemp()
p = write_skm_skd(os.path.join(d, "db"), db)
cli = os.path.join(ROOT, "sketchlib.rust_b200", "bin", "sketchlib")
for args in (["--knn", "50"], ["--knn", "50", "-k", "21"], ["--threads", "1"], ["--threads", str(os.cpu_count())]):
    out = os.path.join(d, "out.txt")
    t = time.time(); r = subprocess.run([cli, "dist", p, "-o", out] + args, capture_output=True, text=True); e = time.time() - t
    print(args, f"{e:.2f} s", "rc", r.returncode, "bytes", os.path.getsize(out), r.stderr.strip()[-60:])
