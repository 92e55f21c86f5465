/*
 * sketcher_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of `sketchlib sketch` for DNA assemblies (FASTA), needed only to
 * make hot-path INPUTS (.skd bit planes) from the reference's test genomes, because
 * the reference binary cannot be built here (no Rust toolchain).  Pinned bit-exactly
 * against the reference's shipped tests/test_files_in/sketches{1,2,3}.skd by
 * tests/test_oracle_golden.py.  Not part of the product.
 *
 * Restates: src/hashing/nthash_iterator.rs (semantics of its own test model: split at
 * non-ACGTU bases and record ends, hash every k-mer, canonical = min(fwd, rc));
 * src/hashing/mod.rs:82-103 (base encoding, srol); src/sketch/mod.rs:132-153, 198-258
 * (bin minima, densify, bit-plane transpose).  The four 64-bit seeds are the published
 * ntHash constants (doi:10.1093/bioinformatics/btw397).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define BBITS 14
static const uint64_t SIGN_MOD = ((uint64_t)1 << 61) - 1; /* src/sketch/mod.rs:36 */

/* index = (ascii >> 1) & 3  =>  A=0, C=1, T/U=2, G=3   (src/hashing/mod.rs:82-84) */
static const uint64_t SEED[4] = {0x3c8bfbb395c60474ULL, 0x3193c18562a02b4cULL, 0x295549f54be24456ULL,
                                 0x20323ed082572324ULL};

static inline int valid_base(uint8_t b) { /* src/hashing/mod.rs:94-97 */
    b |= 0x20;
    return b == 'a' || b == 'c' || b == 'g' || b == 't' || b == 'u';
}
/* rotate_left(1) then swap bits 0 and 33: src/hashing/mod.rs:100-103 */
static inline uint64_t srol(uint64_t v) {
    v = (v << 1) | (v >> 63);
    uint64_t x = (v ^ (v >> 33)) & 1;
    return v ^ (x | (x << 33));
}

/* universal_hash + densify_bin: src/sketch/mod.rs:226-258 */
static uint64_t universal_hash(uint64_t s, uint64_t t) {
    uint64_t x = s * 1009 + t * (1000 * 1000 + 3);
    return (x * 48271 + 11) % (((uint64_t)1 << 31) - 1);
}
static int densify_bin(uint64_t *signs, uint64_t nbins) {
    uint64_t maxval = 0;
    for (uint64_t i = 0; i < nbins; ++i)
        if (signs[i] > maxval) maxval = signs[i];
    if (maxval != UINT64_MAX) return 0;
    for (uint64_t i = 0; i < nbins; ++i) {
        uint64_t j = i, attempts = 0;
        while (signs[j] == UINT64_MAX) {
            j = universal_hash(i, attempts) % nbins;
            attempts += 1;
        }
        signs[i] = signs[j];
    }
    return 1;
}

/* Bin minima of all canonical k-mer hashes of one sample over `nbins` bins, then densify_bin
 * (Sketch::get_signs, src/sketch/mod.rs:132-153; with any nbins: get_signs_no_densify :156-176 + densify_bin, as
 * the inverted index uses them, src/inverted.rs:355-386).  seq: bytes of all records; any non-ACGTU byte is a
 * break, and the caller separates records with a '\n' byte (also a break).  Returns 1 if densified. */
int oracle_bin_minima(const uint8_t *seq, uint64_t len, uint64_t k, uint64_t nbins, int rc, uint64_t *signs) {
    uint64_t bin_size = (SIGN_MOD + nbins - 1) / nbins;
    for (uint64_t b = 0; b < nbins; ++b) signs[b] = UINT64_MAX;
    uint64_t run = 0; /* valid bases since the last break */
    for (uint64_t pos = 0; pos < len; ++pos) {
        if (!valid_base(seq[pos])) {
            run = 0;
            continue;
        }
        run += 1;
        if (run < k) continue;
        /* k-mer = seq[pos-k+1 .. pos]; direct (non-rolling) evaluation of the same function */
        uint64_t fh = 0, rh = 0;
        const uint8_t *km = seq + pos + 1 - k;
        for (uint64_t i = 0; i < k; ++i) fh = srol(fh) ^ SEED[(km[i] >> 1) & 3];
        if (rc) {
            for (uint64_t i = k; i-- > 0;) rh = srol(rh) ^ SEED[((km[i] >> 1) & 3) ^ 2];
            if (rh < fh) fh = rh;
        }
        uint64_t sign = fh % SIGN_MOD;
        uint64_t bin = sign / bin_size;
        if (sign < signs[bin]) signs[bin] = sign;
    }
    return densify_bin(signs, nbins);
}

/*
 * Writes n_k * W * 14 u64 in the .skd per-sample layout (k-major, then 64-bin word, then bit plane;
 * fill_usigs src/sketch/mod.rs:215-223).  Returns 1 if any k needed densifying.
 */
int oracle_sketch_dna(const uint8_t *seq, uint64_t len, const uint64_t *kmers, uint64_t n_k, uint64_t sketch_size,
                      int rc, uint64_t *usigs_out) {
    uint64_t W = (sketch_size + 63) / 64, nbins = W * 64; /* num_bins: src/sketch/mod.rs:49-54 */
    uint64_t *signs = (uint64_t *)malloc(sizeof(uint64_t) * nbins);
    int densified = 0;
    memset(usigs_out, 0, sizeof(uint64_t) * n_k * W * BBITS);
    for (uint64_t ki = 0; ki < n_k; ++ki) {
        densified |= oracle_bin_minima(seq, len, kmers[ki], nbins, rc, signs);
        uint64_t *us = usigs_out + ki * W * BBITS;
        for (uint64_t b = 0; b < nbins; ++b)
            for (int p = 0; p < BBITS; ++p) us[(b / 64) * BBITS + p] |= ((signs[b] >> p) & 1) << (b % 64);
    }
    free(signs);
    return densified;
}

/* fill_usigs alone (src/sketch/mod.rs:215-223): bins (any u64, low 14 bits used) -> bit planes.
 * Used by the synthetic sketch-level generator in tests/bench. */
void oracle_fill_usigs(const uint16_t *bins, uint64_t nbins, uint64_t *usigs_out) {
    memset(usigs_out, 0, sizeof(uint64_t) * (nbins / 64) * BBITS);
    for (uint64_t b = 0; b < nbins; ++b)
        for (int p = 0; p < BBITS; ++p)
            usigs_out[(b / 64) * BBITS + p] |= (uint64_t)((bins[b] >> p) & 1) << (b % 64);
}
