"""Regenerates tests/golden/ from the reference checkout (run in the build container only:
`python tests/golden/make_golden.py`; /root/reference does not exist on the GPU box).

* ref_fixtures/: the binary sketch databases and golden stdout files that the reference's own
  tests hold for the distance path (tests/test_files_in, tests/test_results_correct) -- data
  fixtures, copied byte for byte so the parity tests can run without the checkout.
* rfile_k17-29_s10048.{skd,json}: the 4 test genomes sketched by the oracle's restated sketcher
  with `--k-seq 17,31,4 -s 10000` (reference tests/distance.rs:281-301); input of the
  dists_knn_* / dists_subset goldens.
* rfile_k17-29_s1024.{skd,json}: BASELINE.json configs[0] (default sketch size, k=17..29 step 4).
* inverted.ski, reordered_inv.{ski,skq}: inverted indexes of the same 4 genomes (`inverted build -k 21 -s 10
  --write-skq [--species-names species_names.txt]`, reference tests/inverted.rs:256-266, 360-373) written by this
  repo's C++ builder (the reference ships no .ski); its .skq is checked byte for byte against the reference's
  inverted.skq fixture before anything is written.
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402

REF = os.environ.get("SKETCHLIB_REFERENCE", "/root/reference")
IN = os.path.join(REF, "tests", "test_files_in")
OUT = os.path.join(REF, "tests", "test_results_correct")


def main():
    fx = os.path.join(HERE, "ref_fixtures")
    os.makedirs(fx, exist_ok=True)
    for f in ("sketches1", "sketches2", "sketches3", "legacy_db"):
        for ext in (".skm", ".skd"):
            shutil.copyfile(os.path.join(IN, f + ext), os.path.join(fx, f + ext))
    for f in ("subset.txt", "rfile.txt", "completeness.txt"):
        shutil.copyfile(os.path.join(IN, f), os.path.join(fx, f))
    for f in ("dists_knn_ca", "dists_knn_jaccard", "dists_knn_ani", "dists_subset", "inverted_precluster",
              "inverted_precluster_ani", "inverted_query_count", "inverted_query_any", "inverted_query_all",
              "assembly_sketch_info", "assembly_sketch_full_info", "inverted_sketch_info", "inverted_sketch_full_info",
              "inverted_sketch_info_reorder"):
        shutil.copyfile(os.path.join(OUT, f + ".stdout"), os.path.join(fx, f + ".stdout"))
    shutil.copyfile(os.path.join(OUT, "inverted.skq"), os.path.join(fx, "inverted.skq"))  # 4 samples x 10 u16 bins

    shutil.copyfile(os.path.join(IN, "species_names.txt"), os.path.join(fx, "species_names.txt"))
    import sketchlib.rust_b200.build as b
    cli, _ = b.build_host()
    with tempfile.TemporaryDirectory() as d:
        subprocess.check_call([cli, "inverted", "build", "-o", os.path.join(d, "inverted"), "-k", "21", "-s", "10", "-f",
                               "rfile.txt", "--write-skq", "--quiet"], cwd=IN)
        assert open(os.path.join(d, "inverted.skq"), "rb").read() == open(os.path.join(OUT, "inverted.skq"), "rb").read()
        subprocess.check_call([cli, "inverted", "build", "-o", os.path.join(d, "reordered_inv"), "-k", "21", "-s", "10",
                               "-f", "rfile.txt", "--write-skq", "--species-names", "species_names.txt", "--quiet"], cwd=IN)
        for f in ("inverted.ski", "reordered_inv.ski", "reordered_inv.skq"):
            shutil.copyfile(os.path.join(d, f), os.path.join(fx, f))

    names = [l.split()[0] for l in open(os.path.join(IN, "rfile.txt"))]
    recs = {n: oracle.read_fasta_records(os.path.join(IN, n)) for n in names}
    for tag, sketch_size in (("s10048", 10000), ("s1024", 1000)):
        kmers = list(range(17, 31, 4))  # --k-seq 17,31,4 -> 17,21,25,29
        W = (sketch_size + 63) // 64
        rows = []
        dens = False
        for n in names:
            us, d = oracle.sketch_dna(recs[n], kmers, sketch_size)
            rows.append(us)
            dens |= d
        bins = np.concatenate(rows).astype("<u8")
        base = os.path.join(HERE, f"rfile_k17-29_{tag}")
        bins.tofile(base + ".skd")
        json.dump({"names": names, "kmer_lengths": kmers, "sketchsize64": W, "sketch_size": 64 * W,
                   "kmer_stride": W * 14, "sample_stride": W * 14 * len(kmers), "densified": dens},
                  open(base + ".json", "w"), indent=1)
        print("wrote", base, bins.size * 8, "bytes")


if __name__ == "__main__":
    main()
