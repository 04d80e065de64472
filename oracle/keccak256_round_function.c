/* ORACLE -- TEST INFRASTRUCTURE ONLY (see gl.h header).
 *
 * Sequential CPU restatement of the keccak256 precompile circuit:
 *   keccak256_round_function_entry_point   /root/reference/src/keccak256_round_function/mod.rs:673-794
 *   keccak256_precompile_inner             /root/reference/src/keccak256_round_function/mod.rs:155-670
 *   Keccak256PrecompileCallParams          /root/reference/src/keccak256_round_function/mod.rs:45-98
 *   keccak256_absorb_and_run_permutation   /root/reference/src/keccak256_round_function/mod.rs:796-838
 *   ByteBuffer::{fill_with_bytes, consume} /root/reference/src/keccak256_round_function/buffer/mod.rs:73-170
 *   ConditionalWitnessAllocator            /root/reference/src/storage_application/mod.rs:95-229
 * keccak-f[1600] itself lives in un-vendored boojum (gadgets::keccak256::round_function); it is the standard
 * FIPS-202 permutation -- PINNED: the reference's own tests compare the circuit's digest with sha3::Keccak256 for 10
 * (length, unalignment) cases (mod.rs:1096-1144); tests/test_oracle_keccak.py reproduces them against an independent
 * Keccak-256 and hashlib's sha3 permutation.  Queue hashes remain Poseidon2 (parity unpinned).
 */
#include "oracle.h"
#include <string.h>

static const uint64_t KECCAK_RC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL,
    0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL,
    0x0000000080008009ULL, 0x000000008000000aULL, 0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL,
    0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
static const int KECCAK_ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};

static uint64_t rotl64(uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }

/* A[x + 5y] */
void orc_keccak_f1600(uint64_t A[25]) {
    for (int round = 0; round < 24; round++) {
        uint64_t C[5], D[5], B[25];
        for (int x = 0; x < 5; x++) C[x] = A[x] ^ A[x + 5] ^ A[x + 10] ^ A[x + 15] ^ A[x + 20];
        for (int x = 0; x < 5; x++) D[x] = C[(x + 4) % 5] ^ rotl64(C[(x + 1) % 5], 1);
        for (int i = 0; i < 25; i++) A[i] ^= D[i % 5];
        for (int x = 0; x < 5; x++)
            for (int y = 0; y < 5; y++) B[y + 5 * ((2 * x + 3 * y) % 5)] = rotl64(A[x + 5 * y], KECCAK_ROT[x + 5 * y]);
        for (int y = 0; y < 5; y++)
            for (int x = 0; x < 5; x++) A[x + 5 * y] = B[x + 5 * y] ^ (~B[(x + 1) % 5 + 5 * y] & B[(x + 2) % 5 + 5 * y]);
        A[0] ^= KECCAK_RC[round];
    }
}

/* mod.rs:796-838 on the [i][j][byte] byte layout of the FSM */
static void absorb_and_permute(uint8_t state[200], const uint8_t block[136], uint8_t digest[32]) {
    uint64_t A[25];
    for (int i = 0; i < 5; i++)
        for (int j = 0; j < 5; j++) {
            uint64_t lane = 0;
            for (int b = 0; b < 8; b++) lane |= (uint64_t)state[(i * 5 + j) * 8 + b] << (8 * b);
            const int idx = i + 5 * j;
            if (idx < 17)
                for (int b = 0; b < 8; b++) lane ^= (uint64_t)block[8 * idx + b] << (8 * b);
            A[idx] = lane;
        }
    orc_keccak_f1600(A);
    for (int i = 0; i < 5; i++)
        for (int j = 0; j < 5; j++)
            for (int b = 0; b < 8; b++) state[(i * 5 + j) * 8 + b] = (uint8_t)(A[i + 5 * j] >> (8 * b));
    for (int i = 0; i < 4; i++) memcpy(digest + 8 * i, state + (i * 5 + 0) * 8, 8);
}

/* plain Keccak-256 of a byte string (used by the tests as a second check of the permutation) */
void orc_keccak256(const uint8_t *msg, size_t len, uint8_t digest[32]) {
    uint8_t state[200], block[136];
    memset(state, 0, sizeof state);
    while (len >= 136) { absorb_and_permute(state, msg, digest); msg += 136; len -= 136; }
    memset(block, 0, sizeof block);
    memcpy(block, msg, len);
    block[len] ^= 0x01; block[135] ^= 0x80;
    absorb_and_permute(state, block, digest);
}

size_t orc_keccak_encode_fsm(const zkc_keccak_fsm *f, uint64_t *dst) {
    size_t n = 0;
    dst[n++] = f->read_precompile_call; dst[n++] = f->read_unaligned_words_for_round;
    dst[n++] = f->padding_round; dst[n++] = f->completed;
    for (int i = 0; i < 200; i++) dst[n++] = f->keccak_internal_state[i];
    dst[n++] = f->timestamp_to_use_for_read; dst[n++] = f->timestamp_to_use_for_write;
    dst[n++] = f->input_page; dst[n++] = f->input_memory_byte_offset; dst[n++] = f->input_memory_byte_length;
    dst[n++] = f->output_page; dst[n++] = f->output_word_offset; dst[n++] = f->needs_full_padding_round;
    for (int i = 0; i < 192; i++) dst[n++] = f->buffer_bytes[i];
    dst[n++] = f->buffer_filled;
    n += orc_put_queue_state4(dst + n, &f->log_queue_state);
    memcpy(dst + n, f->memory_queue_state.head, 96); n += 12;
    memcpy(dst + n, f->memory_queue_state.tail, 96); n += 12;
    dst[n++] = f->memory_queue_state.length;
    return n; /* 439 */
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

int orc_keccak256_entry_point(zkc_keccak_closed_form *io, const zkc_log_query *requests, size_t n_requests,
                              const uint32_t *memory_reads, size_t n_reads, size_t limit,
                              const zkc_precompile_options *options, uint64_t *trace, uint64_t *memory_states,
                              size_t *n_memory_states, uint64_t commitment[4], zkc_status *status) {
    zkc_status st = {ZKC_OK, 0, -1, 0, 0};
    const int start = io->start_flag != 0;
    const uint32_t formal_address = options && options->precompile_address ? options->precompile_address : ZKC_KECCAK256_PRECOMPILE_ADDRESS_DEFAULT;
    const uint32_t aux_byte = options && options->aux_byte ? options->aux_byte : ZKC_PRECOMPILE_AUX_BYTE_DEFAULT;
    static const uint64_t zero12[12] = {0};
    if (memcmp(io->initial_log_queue_state.head, zero12, 32) || memcmp(io->initial_memory_queue_state.head, zero12, 96))
        fail(&st, -1, ZKC_KC_CHK_TRIVIAL_HEAD);
    zkc_keccak_fsm s;
    memset(&s, 0, sizeof s);
    if (start) s.read_precompile_call = 1; /* :733-741 */
    else s = io->hidden_fsm_input;
    zkc_queue_state4 rq = start ? io->initial_log_queue_state : io->hidden_fsm_input.log_queue_state;
    zkc_queue_state12 mq = start ? io->initial_memory_queue_state : io->hidden_fsm_input.memory_queue_state;

    /* :196-213 */
    const int cfi = s.read_precompile_call && rq.length == 0;
    if (cfi) { s.read_precompile_call = 0; s.read_unaligned_words_for_round = 0; s.completed = 1; }

    size_t rpos = 0, mpos = 0, n_states = 0;
    for (size_t cyc = 0; cyc < limit; cyc++) {
        const uint32_t flags_in[4] = {s.read_precompile_call, s.read_unaligned_words_for_round, s.padding_round, s.completed};
        /* :260 */
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
        }
        /* :284-318 */
        const uint32_t new_len = call.key[1];
        if (read_call) {
            s.input_memory_byte_offset = call.key[0];
            s.input_memory_byte_length = call.key[1];
            s.output_word_offset = call.key[2];
            s.input_page = call.key[4];
            s.output_page = call.key[5];
            s.needs_full_padding_round = (call.key[1] % ZKC_KECCAK_RATE_BYTES) == 0;
            s.timestamp_to_use_for_read = call.timestamp;
            s.timestamp_to_use_for_write = s.timestamp_to_use_for_read + 1;
        }
        /* :321-349 */
        const int reset_buffer = read_call || s.completed;
        const int read_zero = read_call && new_len == 0;
        const int read_nonzero = read_call && new_len != 0;
        s.read_precompile_call = 0;
        s.read_unaligned_words_for_round = s.read_unaligned_words_for_round || read_nonzero;
        s.padding_round = s.padding_round || read_zero;
        if (reset_buffer) {
            memset(s.buffer_bytes, 0, sizeof s.buffer_bytes); s.buffer_filled = 0;
            memset(s.keccak_internal_state, 0, sizeof s.keccak_internal_state);
        }
        if (trace) {
            for (int i = 0; i < 4; i++) T(ZKC_KC_FLAGS_IN + i, cyc) = flags_in[i];
            uint64_t flat[36];
            orc_log_query_flatten(&call, flat);
            for (int i = 0; i < 36; i++) T(ZKC_KC_CALL_ITEM + i, cyc) = flat[i];
            for (int i = 0; i < 4; i++) T(ZKC_KC_REQ_HEAD + i, cyc) = rq.head[i];
            T(ZKC_KC_REQ_LEN, cyc) = rq.length;
            T(ZKC_KC_PARAMS + 0, cyc) = s.input_page; T(ZKC_KC_PARAMS + 1, cyc) = s.input_memory_byte_offset;
            T(ZKC_KC_PARAMS + 2, cyc) = s.input_memory_byte_length; T(ZKC_KC_PARAMS + 3, cyc) = s.output_page;
            T(ZKC_KC_PARAMS + 4, cyc) = s.output_word_offset; T(ZKC_KC_PARAMS + 5, cyc) = s.needs_full_padding_round;
            T(ZKC_KC_TS_READ, cyc) = s.timestamp_to_use_for_read; T(ZKC_KC_TS_WRITE, cyc) = s.timestamp_to_use_for_write;
            T(ZKC_KC_RESET_BUFFER, cyc) = (uint64_t)reset_buffer; T(ZKC_KC_READ_ZERO_LENGTH, cyc) = (uint64_t)read_zero;
            T(ZKC_KC_READ_NON_ZERO_LENGTH, cyc) = (uint64_t)read_nonzero;
        }
        /* :392-492 */
        for (int q = 0; q < ZKC_KECCAK_MEMORY_QUERIES_PER_CYCLE; q++) {
            const uint32_t aligned = s.input_memory_byte_offset / 32, unal = s.input_memory_byte_offset % 32;
            const uint32_t at_most = 32 - unal;
            const uint32_t meaningful = s.input_memory_byte_length < at_most ? s.input_memory_byte_length : at_most;
            const uint32_t next_filled = s.buffer_filled + meaningful;
            if (next_filled > 255) fail(&st, (int64_t)cyc, ZKC_KC_CHK_BUFFER_OVERFLOW);
            const int enough = !(ZKC_KECCAK_BUFFER_SIZE < next_filled);
            const int should_read = meaningful != 0 && enough && s.read_unaligned_words_for_round;
            uint32_t value[8] = {0};
            if (should_read) {
                if (mpos < n_reads) { memcpy(value, memory_reads + 8 * mpos, 32); mpos++; }
                else fail(&st, (int64_t)cyc, ZKC_KC_CHK_WITNESS_EXHAUSTED);
            }
            zkc_memory_query rqry;
            memset(&rqry, 0, sizeof rqry);
            rqry.timestamp = s.timestamp_to_use_for_read; rqry.memory_page = s.input_page; rqry.index = aligned;
            memcpy(rqry.value, value, 32);
            memory_push(&mq, &rqry, should_read, memory_states, &n_states);
            if (should_read) {
                s.input_memory_byte_offset += meaningful; /* add_no_overflow: a legal ABI never wraps */
                s.input_memory_byte_length -= meaningful;
            }
            const uint32_t to_fill = should_read ? meaningful : 0;
            /* fill_with_bytes(be_bytes, offset = unal, to_fill), buffer/mod.rs:73-136 */
            if (to_fill) {
                uint8_t be[32];
                for (int i = 0; i < 32; i++) be[i] = (uint8_t)(value[7 - i / 4] >> (8 * (3 - i % 4)));
                for (uint32_t idx = 0; idx < 32; idx++) {
                    const uint32_t pos = s.buffer_filled + idx;
                    if (pos >= ZKC_KECCAK_BUFFER_SIZE) break;
                    s.buffer_bytes[pos] = (idx < to_fill && unal + idx < 32) ? be[unal + idx] : 0;
                }
                s.buffer_filled += to_fill;
                if (s.buffer_filled > ZKC_KECCAK_BUFFER_SIZE) fail(&st, (int64_t)cyc, ZKC_KC_CHK_BUFFER_OVERFLOW);
            }
            if (trace) {
                const int b = ZKC_KC_QUERY + q * ZKC_KC_QUERY_STRIDE;
                T(b + 0, cyc) = aligned; T(b + 1, cyc) = unal; T(b + 2, cyc) = meaningful; T(b + 3, cyc) = (uint64_t)should_read;
                for (int i = 0; i < 8; i++) T(b + 4 + i, cyc) = value[i];
                for (int i = 0; i < 12; i++) T(b + 12 + i, cyc) = mq.tail[i];
                T(b + 24, cyc) = mq.length; T(b + 25, cyc) = s.input_memory_byte_offset;
                T(b + 26, cyc) = s.input_memory_byte_length; T(b + 27, cyc) = s.buffer_filled;
            }
        }
        /* :494-583 */
        const int zero_bytes_left = s.input_memory_byte_length == 0;
        const uint32_t currently_filled = s.buffer_filled;
        const int do_one_byte = currently_filled == ZKC_KECCAK_RATE_BYTES - 1;
        uint8_t input[136];
        memcpy(input, s.buffer_bytes, 136); /* consume::<136>(allow_partial = true) */
        memmove(s.buffer_bytes, s.buffer_bytes + 136, ZKC_KECCAK_BUFFER_SIZE - 136);
        memset(s.buffer_bytes + (ZKC_KECCAK_BUFFER_SIZE - 136), 0, 136);
        s.buffer_filled = s.buffer_filled < 136 ? 0 : s.buffer_filled - 136;
        const int buffer_now_empty = s.buffer_filled == 0;
        const int apply_padding = zero_bytes_left && buffer_now_empty && s.read_unaligned_words_for_round && !s.needs_full_padding_round;
        if (apply_padding) {
            if (currently_filled < 135) input[currently_filled] = 0x01;
            input[135] = do_one_byte ? 0x81 : 0x80;
        }
        if (s.padding_round) { memset(input, 0, 136); input[0] = 0x01; input[135] = 0x80; }
        uint8_t squeezed[32];
        absorb_and_permute(s.keccak_internal_state, input, squeezed);
        const int write_result = apply_padding || s.padding_round;
        zkc_memory_query wq;
        memset(&wq, 0, sizeof wq);
        wq.timestamp = s.timestamp_to_use_for_write; wq.memory_page = s.output_page; wq.index = s.output_word_offset; wq.rw_flag = 1;
        for (int l = 0; l < 8; l++) /* UInt256::from_be_bytes */
            wq.value[l] = ((uint32_t)squeezed[28 - 4 * l] << 24) | ((uint32_t)squeezed[29 - 4 * l] << 16) |
                          ((uint32_t)squeezed[30 - 4 * l] << 8) | squeezed[31 - 4 * l];
        memory_push(&mq, &wq, write_result, memory_states, &n_states);
        /* :633-664 */
        const int input_is_empty = rq.length == 0;
        const int nothing_left = write_result && input_is_empty, process_next = write_result && !input_is_empty;
        s.read_precompile_call = (uint32_t)process_next;
        s.completed = s.completed || nothing_left;
        s.padding_round = s.read_unaligned_words_for_round && zero_bytes_left && buffer_now_empty && s.needs_full_padding_round;
        s.read_unaligned_words_for_round = !(s.read_precompile_call || s.padding_round || s.completed);
        if (trace) {
            T(ZKC_KC_ZERO_BYTES_LEFT, cyc) = (uint64_t)zero_bytes_left; T(ZKC_KC_CURRENTLY_FILLED, cyc) = currently_filled;
            T(ZKC_KC_DO_ONE_BYTE_OF_PADDING, cyc) = (uint64_t)do_one_byte; T(ZKC_KC_BUFFER_NOW_EMPTY, cyc) = (uint64_t)buffer_now_empty;
            T(ZKC_KC_APPLY_PADDING, cyc) = (uint64_t)apply_padding;
            for (int i = 0; i < 136; i++) T(ZKC_KC_INPUT + i, cyc) = input[i];
            for (int i = 0; i < 200; i++) T(ZKC_KC_STATE_OUT + i, cyc) = s.keccak_internal_state[i];
            T(ZKC_KC_WRITE_RESULT, cyc) = (uint64_t)write_result;
            for (int i = 0; i < 8; i++) T(ZKC_KC_RESULT + i, cyc) = wq.value[i];
            for (int i = 0; i < 12; i++) T(ZKC_KC_WRITE_TAIL + i, cyc) = mq.tail[i];
            T(ZKC_KC_WRITE_LEN, cyc) = mq.length;
            T(ZKC_KC_FLAGS_OUT + 0, cyc) = s.read_precompile_call; T(ZKC_KC_FLAGS_OUT + 1, cyc) = s.read_unaligned_words_for_round;
            T(ZKC_KC_FLAGS_OUT + 2, cyc) = s.padding_round; T(ZKC_KC_FLAGS_OUT + 3, cyc) = s.completed;
            for (int i = 0; i < 192; i++) T(ZKC_KC_BUFFER_OUT + i, cyc) = s.buffer_bytes[i];
        }
    }
    if (n_memory_states) *n_memory_states = n_states;
    if (rq.length == 0 && memcmp(rq.head, rq.tail, 32)) fail(&st, -1, ZKC_KC_CHK_QUEUE_CONSISTENCY); /* :667 */

    const int done = (int)s.completed;
    zkc_keccak_fsm out = s;
    out.log_queue_state = rq;
    out.memory_queue_state = mq;
    out._pad = 0;
    zkc_queue_state12 obs_out;
    memset(&obs_out, 0, sizeof obs_out);
    if (done) obs_out = mq; /* :764-771 */
    if (options && options->compare_expected) {
        static uint64_t a[439], b[439];
        orc_keccak_encode_fsm(&out, a); orc_keccak_encode_fsm(&io->hidden_fsm_output, b);
        if (memcmp(a, b, sizeof a) || memcmp(&obs_out, &io->final_memory_state, sizeof obs_out) || (io->completion_flag != 0) != done)
            if (st.code == ZKC_OK) st.code = ZKC_ERR_FSM_OUTPUT_MISMATCH;
    }
    uint64_t e_in[34], e_out[25], e_fin[439], e_fout[439];
    size_t n_in = orc_put_queue_state4(e_in, &io->initial_log_queue_state);
    memcpy(e_in + n_in, io->initial_memory_queue_state.head, 96); n_in += 12;
    memcpy(e_in + n_in, io->initial_memory_queue_state.tail, 96); n_in += 12;
    e_in[n_in++] = io->initial_memory_queue_state.length;
    memcpy(e_out, obs_out.head, 96); memcpy(e_out + 12, obs_out.tail, 96); e_out[24] = obs_out.length;
    const size_t n_fin = orc_keccak_encode_fsm(&io->hidden_fsm_input, e_fin);
    const size_t n_fout = orc_keccak_encode_fsm(&out, e_fout);
    io->hidden_fsm_output = out;
    io->final_memory_state = obs_out;
    io->completion_flag = (uint32_t)done;
    orc_closed_form_commitment(start, done, e_in, n_in, e_out, 25, e_fin, n_fin, e_fout, n_fout, commitment);
    if (status) *status = st;
    return st.code;
}
