"""The C++ host layer (sketchlib.rust_b200/csrc/host): .skm reader, helpers, and the `sketchlib dist` CLI.
CPU tests pin the reader against the reference's fixtures and the CLI's host-side error behaviour; GPU tests
run the CLI end to end against the reference's golden stdout files."""
import os
import subprocess

import numpy as np
import pytest

import oracle
from helpers import FIXTURES, fixture_db, load_json_db, synth_db, synth_pair, write_skm_skd


@pytest.fixture(scope="module")
def tools():
    import sketchlib.rust_b200.build as b
    b.build()
    return b.build_host()


def run(cmd, **kw):
    return subprocess.run(cmd, capture_output=True, text=True, **kw)


def parse_info(stdout):
    info = {"samples": [], "completeness": [], "warn": [], "fmt": []}
    for line in stdout.splitlines():
        f = line.split("\t")
        if f[0] == "sample":
            info["samples"].append((f[2], int(f[3])))
        elif f[0] == "completeness":
            info["completeness"].append(float(f[2]))
        elif f[0] in ("warn", "fmt"):
            info[f[0]].append(f[1])
        elif f[0] == "kmer_lengths":
            info["kmer_lengths"] = [int(x) for x in f[1:]]
        else:
            info[f[0]] = f[1]
    return info


@pytest.mark.parametrize("name", ["sketches1", "sketches2", "sketches3", "legacy_db"])
def test_cpp_skm_reader_matches_fixtures(tools, name):
    _, skm_info = tools
    r = run([skm_info, os.path.join(FIXTURES, name + ".skm")])
    assert r.returncode == 0, r.stderr
    got = parse_info(r.stdout)
    ms = oracle.MultiSketch(os.path.join(FIXTURES, name))
    assert int(got["sketch_size"]) == ms.sketch_size and int(got["sketchsize64"]) == ms.sketchsize64
    assert got["kmer_lengths"] == ms.kmer_lengths
    assert int(got["kmer_stride"]) == ms.kmer_stride and int(got["sample_stride"]) == ms.sample_stride
    assert got["samples"] == [(m["name"], m["index"]) for m in ms.sketch_metadata]
    assert got["hash_type"] == "DNA" and got["sketch_version"] == ms.meta["sketch_version"]


def test_cpp_subset_completeness_and_errors(tools, tmp_path):
    _, skm_info = tools
    db = os.path.join(FIXTURES, "sketches1")
    r = run([skm_info, db, "--subset", os.path.join(FIXTURES, "subset.txt")])
    got = parse_info(r.stdout)
    # block_reindex order = order of the subset file; index = position in the .skd (multisketch.rs:187-200)
    assert got["samples"] == [("R6.fa.gz", 2), ("TIGR4.fa.gz", 3), ("14412_3#84.contigs_velvet.fa.gz", 1)]
    comp = tmp_path / "c.txt"
    comp.write_text("R6.fa.gz\t0.8\nTIGR4.fa.gz\t 0.9 \nnot_there\t0.5\nbadline\nR6x\tabc\n")
    got = parse_info(run([skm_info, db, "--completeness", str(comp)]).stdout)
    assert got["completeness"] == [1.0, 1.0, 0.8, 0.9]
    assert any("not found in sketch database (ignored): not_there" in w for w in got["warn"])
    assert any("using default 1.0: 14412_3#82.contigs_velvet.fa.gz, 14412_3#84.contigs_velvet.fa.gz" in w for w in got["warn"])
    assert any("Could not parse completeness value for 'R6x'" in w for w in got["warn"])
    missing = tmp_path / "missing_names.txt"
    missing.write_text("nope\n")
    r = run([skm_info, db, "--subset", str(missing)])
    assert r.returncode == 1 and "Could not find requested sample nope in sketch metadata" in r.stderr


def test_cpp_float_format_is_rust_display(tools):
    _, skm_info = tools
    rng = np.random.default_rng(0)
    vals = np.concatenate([rng.random(150, dtype=np.float32), np.float32([1.0, 0.0, 0.5, 1e-7, 0.008820182, 0.99009615,
                                                                        0.35546875, 3.4e-5, 0.1, 0.2137754])])
    args = [skm_info, os.path.join(FIXTURES, "sketches1")]
    for v in vals:
        args += ["--fmt", repr(float(v))]
    got = parse_info(run(args).stdout)["fmt"]
    assert got == [oracle.fmt_f32(v) for v in vals]


def test_cli_host_side_errors(tools, tmp_path):
    cli, _ = tools
    db = os.path.join(FIXTURES, "sketches1")
    pct = tmp_path / "pct.txt"
    pct.write_text("R6.fa.gz\t85.0\nTIGR4.fa.gz\t0.9\n")
    r = run([cli, "dist", db, "-k", "31", "--ref-completeness-file", str(pct)])
    assert r.returncode == 1 and "[0.0, 1.0]" in r.stderr and "R6.fa.gz: 85" in r.stderr  # tests/distance.rs:554-624
    r = run([cli, "dist", db, "-k", "33"])
    assert r.returncode == 101 and "K-mer size 33 not found in file" in r.stderr
    r = run([cli, "dist", str(tmp_path / "nothing")])
    assert r.returncode == 101 and "Could not read sketch metadata from" in r.stderr
    r = run([cli, "dist", db, "--ani"])
    assert r.returncode == 2 and "-k <KMER>" in r.stderr                     # clap: requires("kmer")
    r = run([cli, "dist"])
    assert r.returncode == 2 and "<REF_DB>" in r.stderr
    r = run([cli, "merge", "x", "y"])
    assert r.returncode == 2 and "only `dist`" in r.stderr
    r = run([cli, "dist", db, "--threads", "0"])
    assert r.returncode == 2


# ---------------------------------------------------------------- GPU: end-to-end CLI
def golden(name):
    return open(os.path.join(FIXTURES, name + ".stdout")).read()


@pytest.mark.gpu
def test_cli_reproduces_reference_goldens(tools, tmp_path):
    cli, _ = tools
    db = write_skm_skd(str(tmp_path / "rfile"), load_json_db("s10048"))
    r = run([cli, "dist", db, "--knn", "1"])
    assert r.returncode == 0, r.stderr
    assert r.stdout == golden("dists_knn_ca")                                   # tests/distance.rs:270-301
    assert run([cli, "dist", db + ".skm", "--knn", "1", "-k", "21"]).stdout == golden("dists_knn_jaccard")
    assert run([cli, "dist", db, "--knn", "1", "-k", "21", "--ani"]).stdout == golden("dists_knn_ani")
    r = run([cli, "dist", db, "--subset", os.path.join(FIXTURES, "subset.txt"), "-v"])
    assert r.stdout == golden("dists_subset")                                   # tests/distance.rs:690-722
    assert "Distances: core/accessory regression" in r.stderr and "sketchlib done in" in r.stderr
    # the reference's own binary fixture, k=21: same values as the precluster goldens
    r = run([cli, "dist", os.path.join(FIXTURES, "sketches3"), "-k", "21", "--knn", "1", "--quiet"])
    assert r.stdout == golden("inverted_precluster") and r.stderr == ""
    out = tmp_path / "o.txt"
    r = run([cli, "dist", os.path.join(FIXTURES, "sketches3"), "-k", "21", "--knn", "1", "--ani", "-o", str(out)])
    assert r.stdout == "" and out.read_text() == golden("inverted_precluster_ani")
    # legacy (v0.1.3) database loads and gives the oracle's distances
    r = run([cli, "dist", os.path.join(FIXTURES, "legacy_db")])
    leg = fixture_db("legacy_db")
    want = oracle.self_dists_all(leg)
    assert r.stdout == f"R6.fa.gz\tTIGR4.fa.gz\t{oracle.fmt_f32(want[0])}\t{oracle.fmt_f32(want[1])}\n"


@pytest.mark.gpu
def test_cli_dense_and_cross_modes(tools, tmp_path):
    cli, _ = tools
    ref, qry = synth_pair(70, 9, 16, [17, 21, 25, 29], seed=60, n_families=3)
    rp, qp = write_skm_skd(str(tmp_path / "ref"), ref), write_skm_skd(str(tmp_path / "qry"), qry)
    # dense self, text = oracle rendered with Rust float formatting
    r = run([cli, "dist", rp, "--threads", "4"])
    want = oracle.self_dists_all(ref)
    lines, d = [], 0
    for i in range(70):
        for j in range(i + 1, 70):
            lines.append(f"ref{i}\tref{j}\t{oracle.fmt_f32(want[2 * d])}\t{oracle.fmt_f32(want[2 * d + 1])}\n")
            d += 1
    got = r.stdout.splitlines(keepends=True)
    assert len(got) == len(lines)
    mism = [k for k, (a, b) in enumerate(zip(got, lines)) if a != b]
    for k in mism:  # core/acc may differ in the last f32 digit only (<= 1e-6)
        ga, wa = got[k].split("\t"), lines[k].split("\t")
        assert ga[:2] == wa[:2] and all(abs(float(x) - float(y)) <= 1e-6 for x, y in zip(ga[2:], wa[2:]))
    assert len(mism) <= len(lines) // 100
    # cross dense: ref-outer, query-inner, 3 columns for -k (tests/distance.rs:391-404, 465-490)
    r = run([cli, "dist", rp, qp, "-k", "21"])
    rows = [l.split("\t") for l in r.stdout.splitlines()]
    assert len(rows) == 70 * 9 and rows[0][:2] == ["ref0", "qry0"] and rows[1][:2] == ["ref0", "qry1"]
    wantx = oracle.cross_dists_all(ref, qry, k_idx=1)
    assert [row[2] for row in rows] == [oracle.fmt_f32(v) for v in wantx]
    # cross kNN: one block of knn rows per query; names are query then ref; knn > n_ref clamps
    r = run([cli, "dist", rp, qp, "--knn", "3"])
    rows = [l.split("\t") for l in r.stdout.splitlines()]
    assert len(rows) == 9 * 3 and all(len(x) == 4 for x in rows) and rows[0][0] == "qry0" and rows[0][1].startswith("ref")
    r = run([cli, "dist", rp, qp, "--knn", "500", "-k", "21"])
    assert len(r.stdout.splitlines()) == 9 * 70 and "higher than number of reference samples=70" in r.stderr
    # completeness files accepted in both positions
    comp = tmp_path / "c.txt"
    comp.write_text("".join(f"ref{i}\t0.9\n" for i in range(70)))
    qc = tmp_path / "qc.txt"
    qc.write_text("".join(f"qry{i}\t0.85\n" for i in range(9)))
    r = run([cli, "dist", rp, qp, "-k", "21", "--ref-completeness-file", str(comp), "--query-completeness-file", str(qc)])
    wantc = oracle.cross_dists_all(ref, qry, k_idx=1, ref_c=np.full(70, 0.9), query_c=np.full(9, 0.85))
    gotc = np.array([float(l.split("\t")[2]) for l in r.stdout.splitlines()], dtype=np.float32)
    assert np.abs(gotc.astype(np.float64) - wantc).max() <= 1e-6 and not np.array_equal(gotc, wantx)
    # self kNN with knn >= n clamps to n-1 with a warning (src/lib.rs:379-382)
    r = run([cli, "dist", qp, "--knn", "50", "-k", "21"])
    assert len(r.stdout.splitlines()) == 9 * 8 and "knn=50 is higher than number of samples=9" in r.stderr


# ---------------------------------------------------------------- `sketchlib sketch` (CPU tool, DNA FASTA)
def _write_fasta(path, records, gz=False, width=60):
    import gzip
    text = "".join(f">{name} some description\n" + "\n".join(seq[i:i + width] for i in range(0, len(seq), width)) + "\n"
                   for name, seq in records)
    if gz:
        with gzip.open(path, "wt") as fh:
            fh.write(text)
    else:
        with open(path, "w") as fh:
            fh.write(text)


def test_sketch_cli_matches_oracle_sketcher(tools, tmp_path):
    cli, skm_info = tools
    rng = np.random.default_rng(11)
    def rand_seq(n):
        return "".join(rng.choice(list("ACGT"), size=n))
    anc = rand_seq(30000)
    mut = list(anc)
    for p in rng.choice(len(mut), size=600, replace=False):
        mut[p] = "ACGT"[("ACGT".index(mut[p]) + 1) % 4]
    a = [("c1", anc[:12000] + "NNNN" + anc[12000:20000].lower()), ("c2", anc[20000:]), ("tiny", "ACGTAC")]
    b = [("c1", "".join(mut)[:15000] + "RYK" + "".join(mut)[15000:])]
    _write_fasta(tmp_path / "a.fa", a)
    _write_fasta(tmp_path / "b.fasta.gz", b, gz=True)
    _write_fasta(tmp_path / "b2.fa", [("extra", rand_seq(5000))])
    rfile = tmp_path / "rfile.txt"
    rfile.write_text(f"sampleA\t{tmp_path / 'a.fa'}\nsampleB\t{tmp_path / 'b.fasta.gz'}\t{tmp_path / 'b2.fa'}\n")
    out = str(tmp_path / "db")
    r = run([cli, "sketch", "-o", out, "--k-seq", "15,27,4", "-s", "200", "-f", str(rfile), "--threads", "2", "--quiet"])
    assert r.returncode == 0, r.stderr
    got = np.fromfile(out + ".skd", dtype="<u8")
    kmers = [15, 19, 23, 27]
    recs_a = [s.encode() for _, s in a]
    recs_b = [b[0][1].encode()] + oracle.read_fasta_records(str(tmp_path / "b2.fa"))
    want = np.concatenate([oracle.sketch_dna(recs_a, kmers, 200)[0], oracle.sketch_dna(recs_b, kmers, 200)[0]])
    assert np.array_equal(got, want)
    ms = oracle.MultiSketch(out).read_sketch_data()       # our .skm is readable by the independent reader
    assert ms.names() == ["sampleA", "sampleB"] and ms.kmer_lengths == kmers and ms.sketch_size == 256
    assert ms.sketch_metadata[0]["non_acgt"] == 4 and ms.sketch_metadata[1]["non_acgt"] == 3
    assert ms.sketch_metadata[0]["seq_length"] == 30000 + 6
    info = parse_info(run([skm_info, out]).stdout)        # ... and by the C++ reader
    assert info["samples"] == [("sampleA", 0), ("sampleB", 1)] and info["kmer_lengths"] == kmers
    # single strand changes the hashes; positional FASTA arguments name samples by file name
    r = run([cli, "sketch", "-o", out + "_ss", "--k-vals", "19", "-s", "200", "--single-strand", str(tmp_path / "a.fa")])
    assert r.returncode == 0
    ss = np.fromfile(out + "_ss.skd", dtype="<u8")
    assert np.array_equal(ss, oracle.sketch_dna(recs_a, [19], 200, rc=False)[0])
    assert oracle.MultiSketch(out + "_ss").names() == ["a.fa"]
    # errors: reads, k too long for every contig, missing file, missing k
    fq = tmp_path / "r.fastq"
    fq.write_text("@r1\nACGT\n+\nIIII\n")
    assert run([cli, "sketch", "-o", out, "--k-vals", "3", str(fq)]).returncode == 101
    _write_fasta(tmp_path / "short.fa", [("s", "ACGTACGT")])
    r = run([cli, "sketch", "-o", out, "--k-vals", "21", str(tmp_path / "short.fa")])
    assert r.returncode == 101 and "K-mer larger than smallest valid sequence" in r.stderr
    assert run([cli, "sketch", "-o", out, "--k-vals", "21", str(tmp_path / "nope.fa")]).returncode == 101
    assert run([cli, "sketch", "-o", out, str(tmp_path / "a.fa")]).returncode == 2


def test_sketch_cli_reproduces_reference_skd():
    ref_in = "/root/reference/tests/test_files_in"
    if not os.path.exists(os.path.join(ref_in, "rfile.txt")):
        pytest.skip("reference checkout not present (GPU box)")
    import tempfile
    import sketchlib.rust_b200.build as b
    cli, _ = b.build_host()
    with tempfile.TemporaryDirectory() as d:
        for name, k, s in (("sketches1", "31", "1000"), ("sketches2", "31", "10000"), ("sketches3", "21", "1000")):
            r = run([cli, "sketch", "-o", os.path.join(d, name), "--k-vals", k, "-s", s, "-f", "rfile.txt", "--threads", "4"],
                    cwd=ref_in)
            assert r.returncode == 0, r.stderr
            assert open(os.path.join(d, name + ".skd"), "rb").read() == open(os.path.join(FIXTURES, name + ".skd"), "rb").read()


@pytest.mark.gpu
def test_full_pipeline_fasta_to_distances(tools, tmp_path):
    """FASTA -> `sketchlib sketch` -> .skm/.skd -> `sketchlib dist` on the GPU, against the oracle end to end
    (oracle sketcher + oracle distances on the same sequences)."""
    cli, _ = tools
    rng = np.random.default_rng(21)
    anc = rng.choice(list("ACGT"), size=60000)
    names, recs = [], {}
    for g in range(6):
        seq = anc.copy()
        pos = rng.choice(len(seq), size=300 * (g + 1), replace=False)
        seq[pos] = rng.choice(list("ACGT"), size=len(pos))
        if g % 2:
            seq = np.concatenate([seq[:20000], seq[26000:]])  # accessory deletion
        s = "".join(seq)
        _write_fasta(tmp_path / f"g{g}.fa", [("chr", s[:35000]), ("plasmid", s[35000:])], gz=(g % 3 == 0))
        names.append(f"g{g}.fa")
        recs[names[-1]] = [s[:35000].encode(), s[35000:].encode()]
    files = [str(tmp_path / n) for n in names]
    out = str(tmp_path / "db")
    assert run([cli, "sketch", "-o", out, "--k-seq", "13,29,4", "-s", "1000", "--threads", "3"] + files).returncode == 0
    kmers = [13, 17, 21, 25, 29]
    bins = np.concatenate([oracle.sketch_dna(recs[n], kmers, 1000)[0] for n in names])
    db = oracle.Db(bins, 6, kmers, 16, names=names)
    r = run([cli, "dist", out])
    assert r.returncode == 0, r.stderr
    want = oracle.self_dists_all(db)
    rows = [l.split("\t") for l in r.stdout.splitlines()]
    assert len(rows) == 15 and [row[:2] for row in rows[:2]] == [["g0.fa", "g1.fa"], ["g0.fa", "g2.fa"]]
    got = np.array([[float(row[2]), float(row[3])] for row in rows], dtype=np.float32).ravel()
    assert np.abs(got.astype(np.float64) - want).max() <= 1e-6
    assert 0 < got[0] < 0.2 and got[1] > 0.01  # related genomes with accessory differences
    r = run([cli, "dist", out, "-k", "21", "--knn", "2"])
    oidx, od0, _ = oracle.self_dists_knn(db, 2, k_idx=2)
    assert [l.split("\t")[2] for l in r.stdout.splitlines()] == [oracle.fmt_f32(v) for v in od0.ravel()]


@pytest.mark.gpu
def test_cli_multi_gpu_knn_equals_single_gpu(tools, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    cli, _ = tools
    ref, qry = synth_pair(700, 90, 16, [17, 21, 25, 29], seed=95, n_families=9)
    ref.names = [f"r{i}" for i in range(700)]
    ref = oracle.Db(ref.bins, 700, ref.kmers, ref.W, names=ref.names)
    rp, qp = write_skm_skd(str(tmp_path / "ref"), ref), write_skm_skd(str(tmp_path / "qry"), qry)
    sub = tmp_path / "subset.txt"
    sub.write_text("".join(f"r{i}\n" for i in range(699, 100, -3)))
    for args in (["--knn", "20"], ["--knn", "20", "-k", "21"], ["--knn", "80", "-k", "21", "--ani"], [qp, "--knn", "7"],
                 [qp, "--knn", "7", "-k", "25"], ["--knn", "5", "--subset", str(sub)]):
        one = run([cli, "dist", rp] + args)
        two = run([cli, "dist", rp] + args + ["--gpus", "2"])
        assert one.returncode == 0 and two.returncode == 0, (args, one.stderr, two.stderr)
        assert one.stdout == two.stdout and len(one.stdout) > 0, args
