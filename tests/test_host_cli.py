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


def test_stale_name_map_entries_are_bounds_checked(tools, tmp_path):
    """The reference's `delete` shrinks sketch_metadata but leaves name_map stale (remove_metadata); Rust panics with
    an index-out-of-bounds on such an entry.  Here: a clean error, never an out-of-bounds access."""
    cli, skm_info = tools
    db = synth_db(6, 2, [15, 19], seed=3)
    nm = {n: i for i, n in enumerate(db.names)}
    nm[db.names[2]] = 1048575           # stale: points far past sketch_metadata
    nm["ghost"] = 6                     # deleted sample still in the map, one past the end
    prefix = write_skm_skd(str(tmp_path / "stale"), db, name_map=nm)
    sub = tmp_path / "sub.txt"
    sub.write_text(f"{db.names[0]}\n{db.names[2]}\n")
    comp = tmp_path / "comp.txt"
    comp.write_text(f"{db.names[2]}\t0.9\nghost\t0.5\n")
    for args in ([skm_info, prefix, "--subset", str(sub)], [skm_info, prefix, "--completeness", str(comp)],
                 [cli, "dist", prefix, "--subset", str(sub)], [cli, "dist", prefix, "--ref-completeness-file", str(comp)]):
        r = run(args)
        assert r.returncode in (1, 101), (args, r.returncode, r.stderr)
        assert "index out of bounds" in r.stderr, r.stderr
    # entries that are in range keep working next to the stale ones
    ok = tmp_path / "ok.txt"
    ok.write_text(f"{db.names[4]}\n{db.names[0]}\n")
    got = parse_info(run([skm_info, prefix, "--subset", str(ok)]).stdout)
    assert got["samples"] == [(db.names[4], 4), (db.names[0], 0)]


def test_completeness_values_parse_like_rust(tools, tmp_path):
    """f64::from_str accepts decimal floats, inf and nan; not hex floats or trailing garbage (strtod took both)."""
    _, skm_info = tools
    db = os.path.join(FIXTURES, "sketches1")
    comp = tmp_path / "c.txt"
    comp.write_text("R6.fa.gz\t0x1p-1\nTIGR4.fa.gz\t5e-1\n14412_3#82.contigs_velvet.fa.gz\t.25\n"
                    "14412_3#84.contigs_velvet.fa.gz\t0.5x\n")
    got = parse_info(run([skm_info, db, "--completeness", str(comp)]).stdout)
    assert got["completeness"] == [0.25, 1.0, 1.0, 0.5]
    assert sum("Could not parse completeness value" in w for w in got["warn"]) == 2
    bad = tmp_path / "nan.txt"
    bad.write_text("R6.fa.gz\tnan\n")
    r = run([skm_info, db, "--completeness", str(bad)])
    assert r.returncode == 1 and "R6.fa.gz: NaN" in r.stderr


def test_python_subset_names_are_whole_lines(tmp_path):
    import sketchlib.rust_b200.host as host
    f = tmp_path / "names.txt"
    f.write_bytes(b"a b\r\n\nlast")
    assert host.read_subset_names(str(f)) == ["a b", "", "last"]


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


@pytest.mark.gpu
def test_cli_cross_knn_streams_reference_shards(tools, tmp_path):
    """Cross --knn with the references loaded and compared shard by shard (loader thread + second context, the
    shard handed over with skd_db_adopt): same text as the run that loads the whole reference database at once,
    with and without completeness files, single-k and core/accessory, and in --subset order."""
    cli, _ = tools
    ref, qry = synth_pair(1500, 23, 16, [17, 21, 25, 29], seed=61, n_families=12)
    rp, qp = write_skm_skd(str(tmp_path / "ref"), ref), write_skm_skd(str(tmp_path / "qry"), qry)
    comp = tmp_path / "c.txt"
    rng = np.random.default_rng(4)
    comp.write_text("".join(f"ref{i}\t{c:.3f}\n" for i, c in enumerate(rng.uniform(0.7, 1.0, size=1500))))
    qc = tmp_path / "qc.txt"
    qc.write_text("".join(f"qry{i}\t0.9\n" for i in range(23)))
    sub = tmp_path / "subset.txt"
    sub.write_text("".join(f"ref{i}\n" for i in range(1499, 200, -2)))
    for extra in (["--knn", "10"], ["--knn", "10", "-k", "21", "--ani"], ["--knn", "64", "-k", "25"],
                  ["--knn", "7", "--ref-completeness-file", str(comp), "--query-completeness-file", str(qc)],
                  ["--knn", "5", "--subset", str(sub)]):
        whole = run([cli, "dist", rp, qp] + extra)
        assert whole.returncode == 0, whole.stderr
        shard = run([cli, "dist", rp, qp, "-v", "--ref-shard-mb", "1"] + extra)   # 1 MB = 128 references per shard
        assert shard.returncode == 0, shard.stderr
        assert "references in shards" in shard.stderr and "references in shards" not in whole.stderr
        assert shard.stdout == whole.stdout and len(whole.stdout.splitlines()) == 23 * int(extra[1])
    # the oracle agrees with the sharded run (values; neighbours apart from exact ties)
    r = run([cli, "dist", rp, qp, "--knn", "10", "-k", "21", "--ref-shard-mb", "2"])
    oidx, od0, _ = oracle.cross_dists_knn(ref, qry, 10, k_idx=1, threads=4)
    assert [l.split("\t")[2] for l in r.stdout.splitlines()] == [oracle.fmt_f32(v) for v in od0.ravel()]


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
def test_cli_multi_gpu_equals_single_gpu(tools, tmp_path):
    """--gpus 2 (two ranks of a peer group; on a one-GPU box both ranks share the device) against --gpus 1: sparse self
    and cross outputs (symmetric partial lists, row split for knn > 64, reference shards) and dense outputs."""
    cli, _ = tools
    ref, qry = synth_pair(700, 90, 16, [17, 21, 25, 29], seed=95, n_families=9)
    ref.names = [f"r{i}" for i in range(700)]
    ref = oracle.Db(ref.bins, 700, ref.kmers, ref.W, names=ref.names)
    rp, qp = write_skm_skd(str(tmp_path / "ref"), ref), write_skm_skd(str(tmp_path / "qry"), qry)
    sub = tmp_path / "subset.txt"
    sub.write_text("".join(f"r{i}\n" for i in range(699, 100, -3)))
    env = dict(os.environ, SKD_PEER_TIMEOUT_S="30")
    for args in (["--knn", "20"], ["--knn", "20", "-k", "21"], ["--knn", "80", "-k", "21", "--ani"], [qp, "--knn", "7"],
                 [qp, "--knn", "7", "-k", "25"], ["--knn", "5", "--subset", str(sub)],
                 [], ["-k", "21"], [qp], [qp, "-k", "25", "--ani"], ["--subset", str(sub)]):
        one = run([cli, "dist", rp] + args)
        two = run([cli, "dist", rp] + args + ["--gpus", "2", "--threads", "3"], env=env)
        assert one.returncode == 0 and two.returncode == 0, (args, one.stderr, two.stderr)
        assert one.stdout == two.stdout and len(one.stdout) > 0, args
    # both devices stream their half of the references in several shards (700 refs, 1 MB = 128 refs per shard)
    one = run([cli, "dist", rp, qp, "--knn", "9"])
    two = run([cli, "dist", rp, qp, "--knn", "9", "--gpus", "2", "--ref-shard-mb", "1", "-v"], env=env)
    assert two.returncode == 0 and two.stdout == one.stdout and "references in shards" in two.stderr
    three = run([cli, "dist", rp, qp, "--knn", "9", "--gpus", "3", "-k", "29"], env=env)
    assert three.returncode == 0 and three.stdout == run([cli, "dist", rp, qp, "--knn", "9", "-k", "29"]).stdout


# ---------------------------------------------------------------- inverted index (.ski / .skq) and precluster

def _roaring_values(blob):
    """Independent reader of the portable RoaringBitmap layout without run containers (cookie 12346)."""
    import struct
    cookie, n = struct.unpack_from("<II", blob, 0)
    assert cookie == 12346
    desc = [struct.unpack_from("<HH", blob, 8 + 4 * i) for i in range(n)]
    offs = struct.unpack_from(f"<{n}I", blob, 8 + 4 * n)
    out = []
    for (key, card_m1), off in zip(desc, offs):
        card = card_m1 + 1
        if card > 4096:
            words = np.frombuffer(blob, dtype="<u8", count=1024, offset=off)
            bits = np.unpackbits(words.view(np.uint8), bitorder="little")
            vals = np.flatnonzero(bits)
            assert len(vals) == card
        else:
            vals = np.frombuffer(blob, dtype="<u2", count=card, offset=off)
        out += [(key << 16) | int(v) for v in vals]
    return out


def _read_ski(path):
    import msgpack
    from oracle import skm
    obj = msgpack.unpackb(skm.snappy_frame_decode(open(path, "rb").read()), raw=False, strict_map_key=False)
    assert len(obj) == 9   # index, n_samples, sample_names, metadata, labels, kmer_size, sketch_version, rc, hash_type
    index = [{k: _roaring_values(v) for k, v in bin_.items()} for bin_ in obj[0]]
    return index, obj[1:]


def test_inverted_build_reproduces_reference_skq_and_query_goldens(tools, tmp_path):
    ref_in = "/root/reference/tests/test_files_in"
    if not os.path.exists(os.path.join(ref_in, "rfile.txt")):
        pytest.skip("reference checkout not present (GPU box)")
    cli, _ = tools
    out = str(tmp_path / "inverted")
    r = run([cli, "inverted", "build", "-o", out, "-v", "-k", "21", "-s", "10", "-f", "rfile.txt", "--write-skq"], cwd=ref_in)
    assert r.returncode == 0, r.stderr
    assert "sketch_size=10\nn_samples=4\nkmer=21\nrc=true\ninverted=true" in r.stderr
    # reference tests/inverted.rs:268-272: the .skq the reference's own build wrote
    assert open(out + ".skq", "rb").read() == open(os.path.join(FIXTURES, "inverted.skq"), "rb").read()
    assert open(out + ".ski", "rb").read() == open(os.path.join(FIXTURES, "inverted.ski"), "rb").read()
    r = run([cli, "inverted", "precluster", "-v", "--count", out + ".ski"])   # tests/inverted.rs:277-285
    assert r.stdout == "Identified 2 prefilter pairs from a max of 6\n"
    for qt, gold in ((None, "inverted_query_count"), ("any-bins", "inverted_query_any"), ("all-bins", "inverted_query_all")):
        cmd = [cli, "inverted", "query", "-v", "-f", "rfile.txt", out + ".ski"] + (["--query-type", qt] if qt else [])
        r = run(cmd, cwd=ref_in)
        assert r.returncode == 0, r.stderr
        want = open(os.path.join(FIXTURES, gold + ".stdout")).read().splitlines()
        got = r.stdout.splitlines()
        assert got[0] == want[0] and sorted(got[1:]) == sorted(want[1:])     # rows are unordered in the reference
    # --species-names reorders the index (tests/inverted.rs:131-166, inverted_sketch_info_reorder.stdout)
    files = ["14412_3#82.contigs_velvet.fa.gz", "14412_3#84.contigs_velvet.fa.gz", "R6.fa.gz", "TIGR4.fa.gz"]
    r = run([cli, "inverted", "build", "-o", out + "_re", "-k", "61", "-s", "63", "--species-names", "species_names.txt"] + files,
            cwd=ref_in)
    assert r.returncode == 0, r.stderr
    _, meta = _read_ski(out + "_re.ski")
    assert meta[1] == ["TIGR4.fa.gz", "R6.fa.gz", files[0], files[1]] and meta[3] == ["Spn", "Spn", "N m", ""]
    assert meta[4:] == [61, "0.3.0", True, "DNA"]


def test_inverted_index_content_and_ski_format(tools, tmp_path):
    """Synthetic genomes: the .ski decoded by an independent MessagePack + roaring reader equals the inverted index of
    the .skq bins; queries agree with numpy; an index loaded back from the .ski gives the same answers."""
    cli, _ = tools
    rng = np.random.default_rng(5)
    anc = rng.choice(list("ACGT"), size=(3, 20000))
    names, rows = [], []
    for g in range(12):
        seq = anc[g % 3].copy()
        pos = rng.choice(seq.size, size=40 * (g // 3 + 1), replace=False)
        seq[pos] = rng.choice(list("ACGT"), size=pos.size)
        _write_fasta(tmp_path / f"g{g}.fa", [("c", "".join(seq))])
        names.append(f"g{g}")
        rows.append(f"g{g}\t{tmp_path / f'g{g}.fa'}")
    (tmp_path / "rfile.txt").write_text("\n".join(rows) + "\n")
    out = str(tmp_path / "inv")
    r = run([cli, "inverted", "build", "-o", out, "-k", "17", "-s", "37", "-f", str(tmp_path / "rfile.txt"), "--write-skq",
             "--threads", "3"])
    assert r.returncode == 0, r.stderr
    skq = np.fromfile(out + ".skq", dtype="<u2").reshape(12, 37)
    # bins = low 16 bits of the densified bin minima (src/inverted.rs:392): oracle sketcher, 37 bins, no transpose
    for g in (0, 7):
        signs = oracle.bin_minima(oracle.read_fasta_records(str(tmp_path / f"g{g}.fa")), 17, 37)
        assert np.array_equal(skq[g], (signs & 0xFFFF).astype(np.uint16))
    index, meta = _read_ski(out + ".ski")
    assert meta[0] == 12 and meta[1] == names and meta[2] is None and meta[3] is None and meta[4] == 17
    assert len(index) == 37
    for b in range(37):
        want = {}
        for g in range(12):
            want.setdefault(int(skq[g, b]), []).append(g)
        assert index[b] == want
    rq = run([cli, "inverted", "query", "-f", str(tmp_path / "rfile.txt"), out + ".ski"])
    counts = np.array([[int(x) for x in l.split("\t")[1:]] for l in rq.stdout.splitlines()[1:]])
    assert np.array_equal(counts, (skq[:, None, :] == skq[None, :, :]).sum(axis=2))
    ra = run([cli, "inverted", "query", "-f", str(tmp_path / "rfile.txt"), out + ".ski", "--query-type", "any-bins"])
    for g, line in enumerate(ra.stdout.splitlines()[1:]):
        assert line.split("\t")[1].split(",") == [names[j] for j in np.flatnonzero(counts[g] > 0)]
    pairs = int(((counts > 0).sum() - 12) // 2)
    assert run([cli, "inverted", "precluster", "--count", out]).stdout == f"Identified {pairs} prefilter pairs from a max of 66\n"
    # > 4096 samples with the same value in a bin: bitmap containers (one tiny genome listed 5000 times under new names)
    (tmp_path / "many.txt").write_text("".join(f"s{i}\t{tmp_path / 'g0.fa'}\n" for i in range(5000)) +
                                       f"other\t{tmp_path / 'g1.fa'}\n")
    r = run([cli, "inverted", "build", "-o", out + "_many", "-k", "17", "-s", "3", "-f", str(tmp_path / "many.txt"), "--threads", "8"])
    assert r.returncode == 0, r.stderr
    index, meta = _read_ski(out + "_many.ski")
    assert meta[0] == 5001 and all(sorted(len(v) for v in b.values())[-1] >= 5000 for b in index)
    assert sorted(v for vals in index[0].values() for v in vals) == list(range(5001))
    r = run([cli, "inverted", "query", out + "_many.ski", str(tmp_path / "g1.fa"), "--query-type", "all-bins"])
    assert r.stdout.splitlines()[1] == "g1.fa\tother"
    # errors
    assert run([cli, "inverted", "precluster", out]).returncode == 2                       # needs --skd or --count
    assert run([cli, "inverted", "query", out + ".ski", "--query-type", "x", str(tmp_path / "g0.fa")]).returncode == 2
    assert run([cli, "inverted", "precluster", "--count", str(tmp_path / "nope.ski")]).returncode == 1
    (tmp_path / "bad.ski").write_bytes(b"\xff\x06\x00\x00sNaPpY\x01\x05\x00\x00" + b"\x00" * 5)
    assert run([cli, "inverted", "precluster", "--count", str(tmp_path / "bad.ski")]).returncode == 1


def _precluster_inputs(tmp_path, ski):
    import shutil
    for f in ("sketches3.skm", "sketches3.skd"):
        shutil.copyfile(os.path.join(FIXTURES, f), tmp_path / f.replace("sketches3", "standard"))
    for ext in (".ski", ".skq"):
        shutil.copyfile(os.path.join(FIXTURES, ski + ext), tmp_path / (ski + ext))


@pytest.mark.gpu
@pytest.mark.parametrize("ski", ["inverted", "reordered_inv"])
def test_cli_inverted_precluster_goldens(tools, tmp_path, ski):
    """reference tests/inverted.rs:244-452: `inverted precluster` over the .ski/.skq pair (same order as the .skd, and
    reordered by --species-names) against the reference's golden stdout files."""
    cli, _ = tools
    _precluster_inputs(tmp_path, ski)
    base = [cli, "inverted", "precluster", "-v", "--skd", "standard", ski + ".ski"]
    for extra, gold in ((["--knn", "1"], "inverted_precluster"), (["--knn", "1", "--ani"], "inverted_precluster_ani"),
                        (["--knn", "50"], "inverted_precluster"), ([], "inverted_precluster")):
        r = run(base + extra, cwd=tmp_path)
        assert r.returncode == 0, r.stderr
        assert sorted(r.stdout.splitlines()) == sorted(golden(gold).splitlines()), (extra, r.stdout)
    assert "knn=50 is higher than number of samples=4" in r.stderr
    # completeness: both index orders give the same text (tests/inverted.rs:525-607), equal to the oracle's values
    comp = tmp_path / "c.txt"
    names = ["14412_3#82.contigs_velvet.fa.gz", "14412_3#84.contigs_velvet.fa.gz", "R6.fa.gz", "TIGR4.fa.gz"]
    cv = [0.7, 0.8, 0.9, 0.95]
    comp.write_text("".join(f"{n}\t{c}\n" for n, c in zip(names, cv)))
    r = run(base + ["--knn", "1", "--ref-completeness-file", str(comp)], cwd=tmp_path)
    assert r.returncode == 0, r.stderr
    db = fixture_db("sketches3")
    oidx, od0, _ = oracle.self_dists_knn(db, 1, k_idx=0, completeness=np.array(cv), cutoff=0.64)
    want = [f"{names[i]}\t{names[int(oidx[i, 0])]}\t{oracle.fmt_f32(od0[i, 0])}" for i in range(4)]
    assert r.stdout.splitlines() == want
    # knn=3 lists only true candidates: no self matches, no padding (tests/inverted.rs:657-701)
    r = run(base + ["--knn", "3"], cwd=tmp_path)
    assert sorted(r.stdout.splitlines()) == sorted(golden("inverted_precluster").splitlines())
    for mode in ("singleton", "bruteforce"):   # every genome has a candidate here: same output (:703-803)
        r = run(base + ["--knn", "1", "--retain-unmatched", mode], cwd=tmp_path)
        assert sorted(r.stdout.splitlines()) == sorted(golden("inverted_precluster").splitlines())
    r = run(base + ["-o", "out.tsv", "--knn", "1", "--threads", "2"], cwd=tmp_path)
    assert r.returncode == 0 and r.stdout == "" and sorted((tmp_path / "out.tsv").read_text().splitlines()) == \
        sorted(golden("inverted_precluster").splitlines())


@pytest.mark.gpu
def test_cli_inverted_precluster_unmatched_and_mismatched_inputs(tools, tmp_path):
    """--retain-unmatched on an index in which no bins are shared (k/s chosen so), and the reference's panics for a
    .ski whose samples or k-mer do not match the .skd (src/distances/mod.rs:429-431, src/lib.rs:748-751)."""
    cli, _ = tools
    _precluster_inputs(tmp_path, "inverted")
    db = fixture_db("sketches3")
    names = db.names
    # an .skq/.ski pair in which every sample has unique bins: written through the C++ builder from tiny unrelated genomes
    rng = np.random.default_rng(9)
    rows = []
    for n in names:
        _write_fasta(tmp_path / (n + ".fa"), [("c", "".join(rng.choice(list("ACGT"), size=3000)))])
        rows.append(f"{n}\t{tmp_path / (n + '.fa')}")
    (tmp_path / "r.txt").write_text("\n".join(rows) + "\n")
    assert run([cli, "inverted", "build", "-o", "lonely", "-k", "21", "-s", "4", "-f", "r.txt", "--write-skq"], cwd=tmp_path).returncode == 0
    assert run([cli, "inverted", "precluster", "--count", "lonely.ski"], cwd=tmp_path).stdout.startswith("Identified 0 ")
    base = [cli, "inverted", "precluster", "--skd", "standard", "lonely.ski", "--knn", "2"]
    assert run(base, cwd=tmp_path).stdout == ""                                  # only padding, which is not printed
    r = run(base + ["--retain-unmatched", "singleton"], cwd=tmp_path)
    assert r.stdout.splitlines() == [f"{n}\t{n}\t0" for n in names]
    r = run(base + ["--retain-unmatched", "bruteforce"], cwd=tmp_path)
    oidx, od0, _ = oracle.self_dists_knn(db, 2, k_idx=0)
    got = [l.split("\t") for l in r.stdout.splitlines()]
    assert [(g[0], g[2]) for g in got] == [(names[i], oracle.fmt_f32(od0[i, t])) for i in range(4) for t in range(2)]
    for g, (i, t) in zip(got, [(i, t) for i in range(4) for t in range(2)]):
        assert g[1] != g[0] and (g[1] == names[int(oidx[i, t])] or g[2] == "1")   # neighbours agree apart from exact ties
    # .ski built at another k: set_k panics
    assert run([cli, "inverted", "build", "-o", "k31", "-k", "31", "-s", "4", "-f", "r.txt", "--write-skq"], cwd=tmp_path).returncode == 0
    r = run([cli, "inverted", "precluster", "--skd", "standard", "k31.ski"], cwd=tmp_path)
    assert r.returncode == 101 and "K-mer size 31 used for .ski not found in .skd" in r.stderr
    # .ski without one of the .skd's samples
    (tmp_path / "r3.txt").write_text("\n".join(rows[:3]) + "\n")
    assert run([cli, "inverted", "build", "-o", "three", "-k", "21", "-s", "4", "-f", "r3.txt", "--write-skq"], cwd=tmp_path).returncode == 0
    r = run([cli, "inverted", "precluster", "--skd", "standard", "three.ski"], cwd=tmp_path)
    assert r.returncode == 101 and "could not be found in the .ski" in r.stderr and names[3] in r.stderr


def test_container_decoders_survive_corrupt_files(tools, tmp_path):
    """Byte flips, truncations and insertions in a .ski (snappy frame + MessagePack + roaring) and a .skm (snappy frame +
    CBOR): the C++ readers must reject or accept them cleanly -- no crash, no hang, no giant allocation from a
    corrupted length field."""
    import shutil
    cli, skm_info = tools
    rng = np.random.default_rng(0)
    for name, cmd in (("inverted.ski", lambda p: [cli, "inverted", "precluster", "--count", p]),
                      ("sketches3.skm", lambda p: [skm_info, p[:-4]])):
        raw = bytearray(open(os.path.join(FIXTURES, name), "rb").read())
        p = str(tmp_path / ("f" + os.path.splitext(name)[1]))
        if name.endswith(".skm"):
            shutil.copyfile(os.path.join(FIXTURES, "sketches3.skd"), p[:-4] + ".skd")
        codes = set()
        for it in range(150):
            m = bytearray(raw)
            if it % 3 == 0:
                for _ in range(int(rng.integers(1, 4))):
                    m[int(rng.integers(10, len(m)))] = int(rng.integers(0, 256))
            elif it % 3 == 1:
                m = m[: int(rng.integers(1, len(m)))]
            else:
                pos = int(rng.integers(10, len(m)))
                m[pos:pos] = bytes(rng.integers(0, 256, size=int(rng.integers(1, 9))).tolist())
            open(p, "wb").write(m)
            r = subprocess.run(cmd(p), capture_output=True, timeout=20)
            assert r.returncode in (0, 1, 2, 101), (name, it, r.returncode, r.stderr[-300:])
            codes.add(r.returncode)
        assert 1 in codes or 101 in codes     # most corruptions are detected


def _matches_snapbox(got, want):
    """snapbox `[..]` matches any run of characters inside a line."""
    import re
    g, w = got.splitlines(), want.splitlines()
    return len(g) == len(w) and all(re.fullmatch(".*".join(re.escape(p) for p in wl.split("[..]")), gl) for gl, wl in zip(g, w))


def test_info_goldens_pin_sketch_metadata_and_index_statistics(tools, tmp_path):
    """`sketchlib info` (src/lib.rs:910-947) on databases made by this build's `sketch` / `inverted build` from the
    reference's genomes, against the reference's golden stdout files (tests/sketch.rs:10-44, tests/inverted.rs:14-50,
    131-166): sequence lengths, base counts, ambiguous bases, strand / read flags, and the index's hashes-per-bin
    statistics over 1000 bins."""
    ref_in = "/root/reference/tests/test_files_in"
    if not os.path.exists(os.path.join(ref_in, "rfile.txt")):
        pytest.skip("reference checkout not present (GPU box)")
    cli, _ = tools
    g82, g84 = (os.path.join(ref_in, f) for f in ("14412_3#82.contigs_velvet.fa.gz", "14412_3#84.contigs_velvet.fa.gz"))
    out = str(tmp_path / "assembly")
    assert run([cli, "sketch", "-o", out, "-v", "-k", "31", g82, g84]).returncode == 0
    r = run([cli, "info", out])
    assert r.returncode == 0 and _matches_snapbox(r.stdout, golden("assembly_sketch_info")), r.stdout
    r = run([cli, "info", "--sample-info", out, "-v"])
    assert _matches_snapbox(r.stdout, golden("assembly_sketch_full_info")), r.stdout
    assert "done in" not in r.stderr                                     # info turns the final message off
    inv = str(tmp_path / "inverted")
    assert run([cli, "inverted", "build", "-o", inv, "-v", "-k", "31", g82, g84]).returncode == 0
    r = run([cli, "info", inv + ".ski", "-v"])
    assert _matches_snapbox(r.stdout, golden("inverted_sketch_info")), r.stdout
    r = run([cli, "info", "--sample-info", inv + ".ski", "-v"])
    assert _matches_snapbox(r.stdout, golden("inverted_sketch_full_info")), r.stdout
    files = [g82, g84, os.path.join(ref_in, "R6.fa.gz"), os.path.join(ref_in, "TIGR4.fa.gz")]
    assert run([cli, "inverted", "build", "-o", inv, "-k", "61", "-s", "63", "--species-names",
                os.path.join(ref_in, "species_names.txt")] + files).returncode == 0
    r = run([cli, "info", inv + ".ski", "--sample-info", "-v"])
    assert _matches_snapbox(r.stdout, golden("inverted_sketch_info_reorder")), r.stdout
    # sample names of positional inputs (read_input_fastas, src/io.rs:20-38): basename for FASTA/FASTQ extensions,
    # the whole argument otherwise
    import shutil, gzip
    odd = tmp_path / "genome.fna"
    with gzip.open(g82, "rb") as fh:
        odd.write_bytes(fh.read())
    assert run([cli, "sketch", "-o", out + "2", "-k", "21", str(odd), g84, "--quiet"]).returncode == 0
    names = [l.split("\t")[0] for l in run([cli, "info", "--sample-info", out + "2"]).stdout.splitlines()[1:3]]
    assert names == [str(odd), "14412_3#84.contigs_velvet.fa.gz"]
    assert run([cli, "info", str(tmp_path / "missing")]).returncode == 101
    assert run([cli, "info"]).returncode == 2
