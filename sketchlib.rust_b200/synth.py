"""Synthetic sketch databases for tests and benchmarks (no network, no Rust sketcher in this image).

Sketch-level generator of SURVEY.md section 8d(ii): per family and k-mer length an ancestor draws its bins
uniformly in [0, 2^14); each member keeps a bin with probability s_k = (1 - a) * exp(-mu * k) and
otherwise redraws it.  Families have geometric sizes (mean `family_mean`).  The bins are then bit-sliced
with the rule of the reference's `fill_usigs` (src/sketch/mod.rs:215-223) into the `.skd` layout
u64[sample][k][w][plane].  This is distribution-matched data, not real genomes; the kernels have no
data-dependent early exit, so throughput does not depend on it -- only the epilogue's break index does.

Uses torch only as an array library (CPU or CUDA); it is not part of the distance path.
"""
import numpy as np
import torch

BBITS = 14


def bit_slice(bins: torch.Tensor) -> torch.Tensor:
    """bins: integer tensor [..., nbins] (nbins % 64 == 0, values < 2^14) -> int64 [..., nbins/64 * 14]
    whose bit pattern is the reference's u64 plane words: word [w*14 + p] has bit b = bit p of bin 64w+b."""
    nbins = bins.shape[-1]
    assert nbins % 64 == 0
    b = bins.to(torch.int64).reshape(*bins.shape[:-1], nbins // 64, 64)
    shifts = torch.arange(64, device=bins.device, dtype=torch.int64)
    planes = []
    for p in range(BBITS):
        bit = (b >> p) & 1
        planes.append((bit << shifts).sum(dim=-1))  # disjoint bits: sum == OR; bit 63 wraps to the sign bit
    out = torch.stack(planes, dim=-1)  # [..., W, 14]
    return out.reshape(*bins.shape[:-1], nbins // 64 * BBITS)


def make_bins(n, sketchsize64, kmers, seed=0, family_mean=50, mu_range=(0.001, 0.03), acc_range=(0.0, 0.3),
              device="cpu", chunk=8192, n_families=None, member_seed=None):
    """Returns int64 tensor [n * K * W * 14] (bit pattern = u64 .skd layout) on `device`.
    `n_families` / `member_seed`: the family ancestors depend only on (seed, n_families, sketch shape); with a
    different member_seed the same families get different members -- queries related to a reference set."""
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    K, nbins = len(kmers), 64 * sketchsize64
    n_fam = int(n_families) if n_families else max(1, int(round(n / family_mean)))
    # geometric-ish family sizes: assign each sample a family with exponentially decaying weights
    w = torch.rand(n_fam, generator=g, device=device).clamp_min(1e-6).log().neg()  # Exp(1) weights
    if member_seed is None:
        fam = torch.multinomial(w, n, replacement=True, generator=g)
        anc = torch.randint(0, 1 << BBITS, (n_fam, K, nbins), generator=g, device=device, dtype=torch.int16)
    else:  # ancestors first (independent of n), then the members from their own stream
        anc = torch.randint(0, 1 << BBITS, (n_fam, K, nbins), generator=g, device=device, dtype=torch.int16)
        g = torch.Generator(device=device)
        g.manual_seed(int(member_seed))
        fam = torch.multinomial(w, n, replacement=True, generator=g)
    mu = torch.empty(n, device=device).uniform_(*mu_range, generator=g)
    acc = torch.empty(n, device=device).uniform_(*acc_range, generator=g)
    kt = torch.tensor(list(kmers), device=device, dtype=torch.float32)
    out = torch.empty((n, K * sketchsize64 * BBITS), dtype=torch.int64, device=device)
    for s0 in range(0, n, chunk):
        s1 = min(n, s0 + chunk)
        keep_p = (1.0 - acc[s0:s1, None]) * torch.exp(-mu[s0:s1, None] * kt[None, :])  # [c, K]
        keep = torch.rand((s1 - s0, K, nbins), generator=g, device=device) < keep_p[:, :, None]
        redraw = torch.randint(0, 1 << BBITS, (s1 - s0, K, nbins), generator=g, device=device, dtype=torch.int16)
        bins = torch.where(keep, anc[fam[s0:s1]], redraw)
        out[s0:s1] = bit_slice(bins).reshape(s1 - s0, -1)
    return out.reshape(-1)


def as_u64_numpy(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().numpy().view(np.uint64)
