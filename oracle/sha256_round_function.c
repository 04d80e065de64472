/* ORACLE -- TEST INFRASTRUCTURE ONLY (see gl.h header).
 *
 * Sequential CPU restatement of the sha256 precompile circuit:
 *   sha256_round_function_entry_point  /root/reference/src/sha256_round_function/mod.rs:343-470
 *   sha256_precompile_inner            /root/reference/src/sha256_round_function/mod.rs:88-340
 *   Sha256PrecompileCallParams         /root/reference/src/sha256_round_function/mod.rs:44-84
 * The compression function lives in un-vendored boojum (gadgets::sha256::round_function_over_uint32); it is the
 * FIPS 180-4 compression.  The reference has NO test for this circuit: pinned here against hashlib.sha256
 * (tests/test_oracle_sha256.py: the digest written to memory equals SHA-256 of the pre-padded message).
 */
#include "oracle.h"
#include <string.h>

static const uint32_t SHA_K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
    0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
    0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
    0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
    0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
    0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
const uint32_t ORC_SHA256_IV[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};

static uint32_t rotr32(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

void orc_sha256_compress(uint32_t state[8], const uint32_t m[16]) {
    uint32_t w[64], a = state[0], b = state[1], c = state[2], d = state[3], e = state[4], f = state[5], g = state[6], h = state[7];
    for (int i = 0; i < 16; i++) w[i] = m[i];
    for (int i = 16; i < 64; i++) {
        const uint32_t s0 = rotr32(w[i - 15], 7) ^ rotr32(w[i - 15], 18) ^ (w[i - 15] >> 3);
        const uint32_t s1 = rotr32(w[i - 2], 17) ^ rotr32(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    for (int i = 0; i < 64; i++) {
        const uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25), ch = (e & f) ^ (~e & g);
        const uint32_t t1 = h + S1 + ch + SHA_K[i] + w[i];
        const uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22), maj = (a & b) ^ (a & c) ^ (b & c);
        const uint32_t t2 = S0 + maj;
        h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    state[0] += a; state[1] += b; state[2] += c; state[3] += d; state[4] += e; state[5] += f; state[6] += g; state[7] += h;
}

size_t orc_sha256_encode_fsm(const zkc_sha256_fsm *f, uint64_t *dst) {
    size_t n = 0;
    dst[n++] = f->read_precompile_call; dst[n++] = f->read_words_for_round; dst[n++] = f->completed;
    for (int i = 0; i < 8; i++) dst[n++] = f->sha256_inner_state[i];
    dst[n++] = f->timestamp_to_use_for_read; dst[n++] = f->timestamp_to_use_for_write;
    dst[n++] = f->input_page; dst[n++] = f->input_offset; dst[n++] = f->output_page; dst[n++] = f->output_offset;
    dst[n++] = f->num_rounds;
    n += orc_put_queue_state4(dst + n, &f->log_queue_state);
    memcpy(dst + n, f->memory_queue_state.head, 96); n += 12;
    memcpy(dst + n, f->memory_queue_state.tail, 96); n += 12;
    dst[n++] = f->memory_queue_state.length;
    return n; /* 52 */
}

static void fail(zkc_status *st, int64_t row, uint32_t bit) {
    st->code = ZKC_ERR_UNSATISFIED;
    st->failed_checks |= bit;
    if (row >= 0 && (st->first_bad_row < 0 || row < st->first_bad_row)) st->first_bad_row = row;
}

static void memory_push(zkc_queue_state12 *q, const zkc_memory_query *mq, int execute, uint64_t *states, size_t *n_states) {
    if (!execute) return;
    uint64_t enc[8];
    orc_memory_query_encode(mq, enc);
    memcpy(q->tail, enc, 64);
    orc_poseidon2_permutation(q->tail);
    q->length++;
    if (states) memcpy(states + 12 * *n_states, q->tail, 96);
    (*n_states)++;
}

#define T(col, r) trace[(size_t)(col) * limit + (r)]

int orc_sha256_entry_point(zkc_sha256_closed_form *io, const zkc_log_query *requests, size_t n_requests,
                           const uint32_t *memory_reads, size_t n_reads, size_t limit, const zkc_precompile_options *options,
                           uint64_t *trace, uint64_t *memory_states, size_t *n_memory_states, uint64_t commitment[4],
                           zkc_status *status) {
    zkc_status st = {ZKC_OK, 0, -1, 0, 0};
    const int start = io->start_flag != 0;
    const uint32_t formal_address = options && options->precompile_address ? options->precompile_address : ZKC_SHA256_PRECOMPILE_ADDRESS_DEFAULT;
    const uint32_t aux_byte = options && options->aux_byte ? options->aux_byte : ZKC_PRECOMPILE_AUX_BYTE_DEFAULT;
    static const uint64_t zero12[12] = {0};
    if (memcmp(io->initial_log_queue_state.head, zero12, 32) || memcmp(io->initial_memory_queue_state.head, zero12, 96))
        fail(&st, -1, ZKC_KC_CHK_TRIVIAL_HEAD);
    zkc_sha256_fsm s;
    memset(&s, 0, sizeof s);
    if (start) { s.read_precompile_call = 1; memcpy(s.sha256_inner_state, ORC_SHA256_IV, 32); } /* mod.rs:411-419, input.rs:36-50 */
    else s = io->hidden_fsm_input;
    zkc_queue_state4 rq = start ? io->initial_log_queue_state : io->hidden_fsm_input.log_queue_state;
    zkc_queue_state12 mq = start ? io->initial_memory_queue_state : io->hidden_fsm_input.memory_queue_state;
    const int cfi = s.read_precompile_call && rq.length == 0; /* :121-135 */
    if (cfi) { s.read_precompile_call = 0; s.read_words_for_round = 0; s.completed = 1; }

    size_t rpos = 0, mpos = 0, n_states = 0;
    for (size_t cyc = 0; cyc < limit; cyc++) {
        const uint32_t flags_in[3] = {s.read_precompile_call, s.read_words_for_round, s.completed};
        zkc_log_query call;
        memset(&call, 0, sizeof call);
        const int read_call = (int)s.read_precompile_call;
        if (read_call) {
            if (rpos < n_requests) call = requests[rpos++];
            else fail(&st, (int64_t)cyc, ZKC_KC_CHK_WITNESS_EXHAUSTED);
            uint64_t enc[20];
            orc_log_query_encode(&call, enc);
            orc_log_queue_absorb(rq.head, enc, NULL);
            rq.length--;
            if (ZKC_LQ_AUX(call.flags) != aux_byte) fail(&st, (int64_t)cyc, ZKC_KC_CHK_AUX_BYTE);
            if (call.address[0] != formal_address || call.address[1] || call.address[2] || call.address[3] || call.address[4])
                fail(&st, (int64_t)cyc, ZKC_KC_CHK_ADDRESS);
            s.input_offset = call.key[0]; s.output_offset = call.key[2]; s.input_page = call.key[4];
            s.output_page = call.key[5]; s.num_rounds = call.key[6];
            s.timestamp_to_use_for_read = call.timestamp;
            s.timestamp_to_use_for_write = call.timestamp + 1;
            if (s.num_rounds == 0) fail(&st, (int64_t)cyc, ZKC_SH_CHK_ZERO_ROUNDS);
        }
        const int reset_buffer = read_call || s.completed; /* :204 */
        s.read_words_for_round = read_call || s.read_words_for_round;
        s.read_precompile_call = 0;
        const int should_read = s.num_rounds != 0; /* :213-215 */
        if (trace) {
            for (int i = 0; i < 3; i++) T(ZKC_SH_FLAGS_IN + i, cyc) = flags_in[i];
            uint64_t flat[36];
            orc_log_query_flatten(&call, flat);
            for (int i = 0; i < 36; i++) T(ZKC_SH_CALL_ITEM + i, cyc) = flat[i];
            for (int i = 0; i < 4; i++) T(ZKC_SH_REQ_HEAD + i, cyc) = rq.head[i];
            T(ZKC_SH_REQ_LEN, cyc) = rq.length;
            T(ZKC_SH_PARAMS + 0, cyc) = s.input_page; T(ZKC_SH_PARAMS + 1, cyc) = s.input_offset; T(ZKC_SH_PARAMS + 2, cyc) = s.output_page;
            T(ZKC_SH_PARAMS + 3, cyc) = s.output_offset; T(ZKC_SH_PARAMS + 4, cyc) = s.num_rounds;
            T(ZKC_SH_TS_READ, cyc) = s.timestamp_to_use_for_read; T(ZKC_SH_TS_WRITE, cyc) = s.timestamp_to_use_for_write;
            T(ZKC_SH_RESET_BUFFER, cyc) = (uint64_t)reset_buffer; T(ZKC_SH_SHOULD_READ, cyc) = (uint64_t)should_read;
        }
        uint32_t m[16];
        for (int q = 0; q < 2; q++) {
            uint32_t value[8] = {0};
            if (should_read) {
                if (mpos < n_reads) { memcpy(value, memory_reads + 8 * mpos, 32); mpos++; }
                else fail(&st, (int64_t)cyc, ZKC_KC_CHK_WITNESS_EXHAUSTED);
            }
            zkc_memory_query rqry;
            memset(&rqry, 0, sizeof rqry);
            rqry.timestamp = s.timestamp_to_use_for_read; rqry.memory_page = s.input_page; rqry.index = s.input_offset;
            memcpy(rqry.value, value, 32);
            if (s.read_words_for_round) s.input_offset = s.input_offset + 1; /* :233-244 */
            memory_push(&mq, &rqry, should_read, memory_states, &n_states);
            for (int i = 0; i < 8; i++) m[8 * q + i] = value[7 - i]; /* BE words of the BE memory word, :250-254 */
            if (trace) {
                const int b = ZKC_SH_QUERY + q * ZKC_SH_QUERY_STRIDE;
                for (int i = 0; i < 8; i++) T(b + i, cyc) = value[i];
                for (int i = 0; i < 12; i++) T(b + 8 + i, cyc) = mq.tail[i];
                T(b + 20, cyc) = mq.length; T(b + 21, cyc) = s.input_offset;
            }
        }
        if (s.read_words_for_round) s.num_rounds = s.num_rounds - 1; /* :257-268 (u32 wrap only on an illegal num_rounds = 0) */
        uint32_t cur[8];
        memcpy(cur, reset_buffer ? ORC_SHA256_IV : s.sha256_inner_state, 32); /* :271-278 */
        if (trace) for (int i = 0; i < 8; i++) T(ZKC_SH_STATE_IN + i, cyc) = cur[i];
        orc_sha256_compress(cur, m);
        memcpy(s.sha256_inner_state, cur, 32);
        const int no_rounds_left = s.num_rounds == 0;
        const int write_result = s.read_words_for_round && no_rounds_left;
        zkc_memory_query wq;
        memset(&wq, 0, sizeof wq);
        wq.timestamp = s.timestamp_to_use_for_write; wq.memory_page = s.output_page; wq.index = s.output_offset; wq.rw_flag = 1;
        for (int k = 0; k < 8; k++) wq.value[7 - k] = cur[k]; /* :290-299: H0 is the most significant limb */
        memory_push(&mq, &wq, write_result, memory_states, &n_states);
        const int input_is_empty = rq.length == 0;
        const int nothing_left = write_result && input_is_empty, process_next = write_result && !input_is_empty;
        s.read_precompile_call = (uint32_t)process_next;
        s.completed = s.completed || nothing_left;
        s.read_words_for_round = !(s.read_precompile_call || s.completed);
        if (trace) {
            for (int i = 0; i < 16; i++) T(ZKC_SH_MESSAGE + i, cyc) = m[i];
            T(ZKC_SH_NUM_ROUNDS, cyc) = s.num_rounds;
            for (int i = 0; i < 8; i++) { T(ZKC_SH_STATE_OUT + i, cyc) = cur[i]; T(ZKC_SH_RESULT + i, cyc) = wq.value[i]; }
            T(ZKC_SH_WRITE_RESULT, cyc) = (uint64_t)write_result;
            for (int i = 0; i < 12; i++) T(ZKC_SH_WRITE_TAIL + i, cyc) = mq.tail[i];
            T(ZKC_SH_WRITE_LEN, cyc) = mq.length;
            T(ZKC_SH_FLAGS_OUT + 0, cyc) = s.read_precompile_call; T(ZKC_SH_FLAGS_OUT + 1, cyc) = s.read_words_for_round;
            T(ZKC_SH_FLAGS_OUT + 2, cyc) = s.completed;
        }
    }
    if (n_memory_states) *n_memory_states = n_states;
    if (rq.length == 0 && memcmp(rq.head, rq.tail, 32)) fail(&st, -1, ZKC_KC_CHK_QUEUE_CONSISTENCY);
    const int done = (int)s.completed;
    zkc_sha256_fsm out = s;
    out.log_queue_state = rq;
    out.memory_queue_state = mq;
    out._pad = 0;
    zkc_queue_state12 obs_out;
    memset(&obs_out, 0, sizeof obs_out);
    if (done) obs_out = mq;
    if (options && options->compare_expected) {
        uint64_t a[52], b[52];
        orc_sha256_encode_fsm(&out, a); orc_sha256_encode_fsm(&io->hidden_fsm_output, b);
        if (memcmp(a, b, sizeof a) || memcmp(&obs_out, &io->final_memory_state, sizeof obs_out) || (io->completion_flag != 0) != done)
            if (st.code == ZKC_OK) st.code = ZKC_ERR_FSM_OUTPUT_MISMATCH;
    }
    uint64_t e_in[34], e_out[25], e_fin[52], e_fout[52];
    size_t n_in = orc_put_queue_state4(e_in, &io->initial_log_queue_state);
    memcpy(e_in + n_in, io->initial_memory_queue_state.head, 96); n_in += 12;
    memcpy(e_in + n_in, io->initial_memory_queue_state.tail, 96); n_in += 12;
    e_in[n_in++] = io->initial_memory_queue_state.length;
    memcpy(e_out, obs_out.head, 96); memcpy(e_out + 12, obs_out.tail, 96); e_out[24] = obs_out.length;
    const size_t n_fin = orc_sha256_encode_fsm(&io->hidden_fsm_input, e_fin);
    const size_t n_fout = orc_sha256_encode_fsm(&out, e_fout);
    io->hidden_fsm_output = out;
    io->final_memory_state = obs_out;
    io->completion_flag = (uint32_t)done;
    orc_closed_form_commitment(start, done, e_in, n_in, e_out, 25, e_fin, n_fin, e_fout, n_fout, commitment);
    if (status) *status = st;
    return st.code;
}
