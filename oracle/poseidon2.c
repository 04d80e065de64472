/* ORACLE -- TEST INFRASTRUCTURE ONLY (see gl.h header).
 *
 * Poseidon2 over Goldilocks, state width 12, rate 8, capacity 4: the round function `R` every
 * reference entry point is generic over (CircuitRoundFunction<F, 8, 12, 4>,
 * /root/reference/src/ram_permutation/mod.rs:34; instantiated as
 * boojum::implementations::poseidon2::Poseidon2Goldilocks at :411,:522).
 *
 * PARITY UNPINNED: the permutation lives in the un-vendored dependency `boojum`
 * (git matter-labs/era-boojum, branch main, no lockfile -- /root/reference/Cargo.toml:19) and no
 * reference test asserts a Poseidon2 output.  This file restates the published construction
 * (SURVEY.md 8c): x^7 S-box, 4 + 22 + 4 rounds, one external-matrix multiplication before round 0,
 * external matrix circ(2*M4, M4, M4), inner matrix J + diag(2^s_i), round constants = the
 * Poseidon Goldilocks table regenerated from ChaCha8Rng::seed_from_u64(0) (checksum pinned in
 * tests/test_oracle_poseidon2.py), partial-round constant i = lane 0 of row 4 + i.
 */
#include "oracle.h"
#include <string.h>

/* ---- constant generation: PCG32 seed expansion -> ChaCha8 -> uniform sampling in [0, p) ---- */
static uint32_t rotl32(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
static uint32_t rotr32(uint32_t x, int n) { n &= 31; return n ? (x >> n) | (x << (32 - n)) : x; }

static void chacha8_block(const uint32_t key[8], uint64_t counter, uint32_t out[16]) {
    uint32_t st[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
    uint32_t x[16];
    memcpy(st + 4, key, 32);
    st[12] = (uint32_t)counter; st[13] = (uint32_t)(counter >> 32); st[14] = 0; st[15] = 0;
    memcpy(x, st, sizeof x);
#define QR(a, b, c, d)                                                    \
    x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 16); x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 12); \
    x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 8);  x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 7);
    for (int r = 0; r < 4; r++) {
        QR(0, 4, 8, 12) QR(1, 5, 9, 13) QR(2, 6, 10, 14) QR(3, 7, 11, 15)
        QR(0, 5, 10, 15) QR(1, 6, 11, 12) QR(2, 7, 8, 13) QR(3, 4, 9, 14)
    }
#undef QR
    for (int i = 0; i < 16; i++) out[i] = x[i] + st[i];
}

void orc_poseidon2_constants(uint64_t out[ORC_P2_NUM_CONSTANTS]) {
    uint32_t key[8], buf[16];
    uint64_t s = 0, counter = 0;
    for (int i = 0; i < 8; i++) {
        s = s * 6364136223846793005ULL + 11634580027462260723ULL;
        uint32_t xs = (uint32_t)(((s >> 18) ^ s) >> 27);
        key[i] = rotr32(xs, (int)(s >> 59));
    }
    const uint64_t range = GL_P;
    const uint64_t zone = (range << __builtin_clzll(range)) - 1;
    int have = 0, pos = 0, n = 0;
    while (n < ORC_P2_NUM_CONSTANTS) {
        if (have - pos < 2) { chacha8_block(key, counter++, buf); have = 16; pos = 0; }
        uint64_t v = (uint64_t)buf[pos] | ((uint64_t)buf[pos + 1] << 32);
        pos += 2;
        u128 m = (u128)v * range;
        if ((uint64_t)m <= zone) out[n++] = (uint64_t)(m >> 64);
    }
}

/* ---- permutation ---- */
static uint64_t RC[ORC_P2_NUM_CONSTANTS];
static int rc_ready = 0;
static const int INNER_SHIFTS[12] = {4, 14, 11, 8, 0, 5, 2, 9, 13, 6, 3, 12};

static void ensure_rc(void) {
    if (!rc_ready) { orc_poseidon2_constants(RC); rc_ready = 1; }
}

static inline uint64_t sbox7(uint64_t x) {
    uint64_t x2 = gl_mul(x, x), x3 = gl_mul(x2, x), x4 = gl_mul(x2, x2);
    return gl_mul(x3, x4);
}

/* M4 = [[5,7,1,3],[4,6,1,1],[1,3,5,7],[1,1,4,6]] on each 4-block, then circ(2,1,1) over the blocks.
 * Coefficients are tiny, so everything is accumulated on 128-bit integers and reduced once per lane. */
static void external_matrix(uint64_t s[12]) {
    u128 t[12];
    for (int b = 0; b < 3; b++) {
        const u128 x0 = s[4 * b], x1 = s[4 * b + 1], x2 = s[4 * b + 2], x3 = s[4 * b + 3];
        t[4 * b + 0] = 5 * x0 + 7 * x1 + x2 + 3 * x3;
        t[4 * b + 1] = 4 * x0 + 6 * x1 + x2 + x3;
        t[4 * b + 2] = x0 + 3 * x1 + 5 * x2 + 7 * x3;
        t[4 * b + 3] = x0 + x1 + 4 * x2 + 6 * x3;
    }
    for (int i = 0; i < 4; i++) {
        const u128 sum = t[i] + t[4 + i] + t[8 + i];
        for (int b = 0; b < 3; b++) s[4 * b + i] = gl_reduce128(t[4 * b + i] + sum);
    }
}

static void inner_matrix(uint64_t s[12]) {
    u128 sum = 0;
    for (int i = 0; i < 12; i++) sum += s[i];
    for (int i = 0; i < 12; i++) s[i] = gl_reduce128(((u128)s[i] << INNER_SHIFTS[i]) + sum);
}

static void full_round(uint64_t s[12], int round) {
    for (int i = 0; i < 12; i++) s[i] = sbox7(gl_add(s[i], RC[12 * round + i]));
    external_matrix(s);
}

void orc_poseidon2_permutation(uint64_t s[12]) {
    ensure_rc();
    int round = 0;
    external_matrix(s);
    for (int i = 0; i < 4; i++) full_round(s, round++);
    for (int i = 0; i < 22; i++) {
        s[0] = sbox7(gl_add(s[0], RC[12 * round]));
        inner_matrix(s);
        round++;
    }
    for (int i = 0; i < 4; i++) full_round(s, round++);
}

/* R::create_empty_state + R::apply_length_specialization (/root/reference/src/utils.rs:31-33,
 * fsm_input_output/mod.rs:297-299).  [from memory of boojum: the length goes to the LAST state
 * element; unpinned] */
void orc_sponge_init(uint64_t s[12], uint64_t length) {
    memset(s, 0, 12 * sizeof(uint64_t));
    s[11] = gl_canon(length);
}

/* absorb_with_replacement(chunk, capacity) + compute_round_function */
void orc_sponge_absorb8(uint64_t s[12], const uint64_t chunk[8]) {
    memcpy(s, chunk, 8 * sizeof(uint64_t));
    orc_poseidon2_permutation(s);
}

/* commit_encoding, /root/reference/src/fsm_input_output/mod.rs:281-326 */
void orc_commit_encoding(const uint64_t *input, size_t n, uint64_t out[4]) {
    uint64_t s[12], chunk[8];
    orc_sponge_init(s, n);
    for (size_t off = 0; off < n; off += 8) {
        for (size_t j = 0; j < 8; j++) chunk[j] = off + j < n ? input[off + j] : 0;
        orc_sponge_absorb8(s, chunk);
    }
    memcpy(out, s, 4 * sizeof(uint64_t));
}

/* ClosedFormInputCompactForm::from_full_form + commit_variable_length_encodable_item of the
 * compact form, /root/reference/src/fsm_input_output/mod.rs:178-255 and e.g.
 * ram_permutation/mod.rs:200-203.  Inputs are the flattened var-length encodings. */
void orc_closed_form_commitment(int start_flag, int completion_flag,
                                const uint64_t *obs_in, size_t n_obs_in,
                                const uint64_t *obs_out, size_t n_obs_out,
                                const uint64_t *fsm_in, size_t n_fsm_in,
                                const uint64_t *fsm_out, size_t n_fsm_out,
                                uint64_t out[4]) {
    uint64_t compact[18];
    compact[0] = start_flag ? 1 : 0;
    compact[1] = completion_flag ? 1 : 0;
    orc_commit_encoding(obs_in, n_obs_in, compact + 2);
    orc_commit_encoding(obs_out, n_obs_out, compact + 6);
    orc_commit_encoding(fsm_in, n_fsm_in, compact + 10);
    orc_commit_encoding(fsm_out, n_fsm_out, compact + 14);
    if (start_flag) memset(compact + 10, 0, 32);       /* hidden input masked at start */
    if (!completion_flag) memset(compact + 6, 0, 32);  /* observable output only when done */
    if (completion_flag) memset(compact + 14, 0, 32);  /* hidden output masked when done */
    orc_commit_encoding(compact, 18, out);
}

/* produce_fs_challenges, /root/reference/src/utils.rs:12-78.  tails are `tw` wide (12 for the
 * full-state RAM queues, 4 for log queues); result[rep][0] = 1, result[rep][1..=enc] squeezed. */
void orc_produce_fs_challenges(const uint64_t *unsorted_tail, uint32_t unsorted_len,
                               const uint64_t *sorted_tail, uint32_t sorted_len, int tw,
                               int num_challenges, uint64_t *result /* [2][num_challenges] */) {
    uint64_t in[26], s[12], chunk[8];
    int n = 0;
    for (int i = 0; i < tw; i++) in[n++] = unsorted_tail[i];
    in[n++] = unsorted_len;
    for (int i = 0; i < tw; i++) in[n++] = sorted_tail[i];
    in[n++] = sorted_len;
    orc_sponge_init(s, (uint64_t)n);
    for (int off = 0; off < n; off += 8) {
        for (int j = 0; j < 8; j++) chunk[j] = off + j < n ? in[off + j] : 0;
        orc_sponge_absorb8(s, chunk);
    }
    int can_take = 8;
    for (int rep = 0; rep < 2; rep++) {
        result[rep * num_challenges] = 1;
        for (int k = 1; k < num_challenges; k++) {
            if (can_take == 0) { orc_poseidon2_permutation(s); can_take = 8; }
            result[rep * num_challenges + k] = s[8 - can_take];
            can_take--;
        }
    }
}
