/* ORACLE -- TEST INFRASTRUCTURE ONLY (see gl.h header).
 *
 * Sequential CPU restatement of the code decommitter (bytecode -> memory, SHA-256 of the code checked against the hash):
 *   unpack_code_into_memory_entry_point   /root/reference/src/code_unpacker_sha256/mod.rs:33-148
 *   unpack_code_into_memory_inner         /root/reference/src/code_unpacker_sha256/mod.rs:150-453
 *   CodeDecommittmentFSM                  /root/reference/src/code_unpacker_sha256/input.rs:27-38
 *   ConditionalWitnessAllocator           /root/reference/src/storage_application/mod.rs:95-229
 * ContractCodeSha256::VERSION_BYTE (= 1) is from the un-vendored zkevm_opcode_defs.
 * Pinning: pinned by the reference's own vector (test_code_unpacker_inner, mod.rs:472-700: one request, 33 words, limit 40,
 * extracted to tests/golden/code_unpacker_vector.json): every enforcement holds -- the versioned-hash format, word order,
 * padding and the hash comparison among them -- the requests queue ends empty and the memory queue equals the queue of the 33
 * writes the test rebuilds; the SHA-256 is additionally pinned against hashlib.  Queue-state / commitment VALUES: PARITY UNPINNED
 * (Poseidon2, see poseidon2.c).
 */
#include "oracle.h"
#include <string.h>

extern const uint32_t ORC_SHA256_IV[8];

static void fail(zkc_status *st, int64_t row, uint32_t bit) {
    st->code = ZKC_ERR_UNSATISFIED;
    st->failed_checks |= bit;
    if (row >= 0 && (st->first_bad_row < 0 || row < st->first_bad_row)) st->first_bad_row = row;
}

static size_t put_queue_state12(uint64_t *dst, const zkc_queue_state12 *s) {
    memcpy(dst, s->head, 96);
    memcpy(dst + 12, s->tail, 96);
    dst[24] = s->length;
    return 25;
}

/* CSVarLengthEncodable order of CodeDecommitterFSMInputOutput (input.rs:70-74) over CodeDecommittmentFSM (:27-38) */
size_t orc_code_unpacker_encode_fsm(const zkc_code_unpacker_fsm *f, uint64_t *dst) {
    const zkc_code_decommittment_fsm *s = &f->internal_fsm;
    size_t n = 0;
    for (int i = 0; i < 8; i++) dst[n++] = s->sha256_inner_state[i];
    for (int i = 0; i < 8; i++) dst[n++] = s->hash_to_compare_against[i];
    dst[n++] = s->current_index; dst[n++] = s->current_page; dst[n++] = s->timestamp;
    dst[n++] = s->num_rounds_left; dst[n++] = s->length_in_bits;
    dst[n++] = s->state_get_from_queue; dst[n++] = s->state_decommit; dst[n++] = s->finished;
    n += put_queue_state12(dst + n, &f->decommittment_requests_queue_state);
    n += put_queue_state12(dst + n, &f->memory_queue_state);
    return n; /* 74 */
}

static void memory_push(zkc_queue_state12 *q, uint32_t ts, uint32_t page, uint32_t index, const uint32_t value[8], int execute,
                        uint64_t *states, size_t *n_states) {
    if (!execute) return;
    zkc_memory_query mq;
    memset(&mq, 0, sizeof mq);
    mq.timestamp = ts; mq.memory_page = page; mq.index = index; mq.rw_flag = 1;
    memcpy(mq.value, value, 32);
    uint64_t enc[8];
    orc_memory_query_encode(&mq, enc);
    memcpy(q->tail, enc, 64);
    orc_poseidon2_permutation(q->tail);
    q->length++;
    if (states) memcpy(states + 12 * *n_states, q->tail, 96);
    (*n_states)++;
}

#define T(col, r) trace[(size_t)(col) * limit + (r)]

int orc_code_unpacker_entry_point(zkc_code_unpacker_closed_form *io, const zkc_decommit_query *requests, size_t n_requests,
                                  const uint32_t *code_words, size_t n_code_words, size_t limit, const zkc_sorter_options *options,
                                  uint64_t *trace, uint64_t *memory_states, size_t *n_memory_states, uint64_t commitment[4],
                                  zkc_status *status) {
    zkc_status st = {ZKC_OK, 0, -1, 0, 0};
    const int start = io->start_flag != 0;
    const zkc_code_unpacker_fsm *fin = &io->hidden_fsm_input;
    zkc_queue_state12 rq = start ? io->sorted_requests_queue_initial_state : fin->decommittment_requests_queue_state; /* :59-69 */
    zkc_queue_state12 mq = start ? io->memory_queue_initial_state : fin->memory_queue_state;                          /* :77-84 */
    zkc_code_decommittment_fsm s;
    memset(&s, 0, sizeof s);
    if (start) s.state_get_from_queue = 1; /* :86-95 */
    else s = fin->internal_fsm;
    s.state_get_from_queue &= 1; s.state_decommit &= 1; s.finished &= 1; s.num_rounds_left &= 0xFFFF;

    size_t rpos = 0, wpos = 0, n_states = 0;
    for (size_t cyc = 0; cyc < limit; cyc++) {
        const zkc_code_decommittment_fsm in = s;
        zkc_decommit_query req;
        memset(&req, 0, sizeof req);
        if (s.state_get_from_queue && (rq.length == 0 || rpos >= n_requests)) {
            /* a pop from the empty queue / an exhausted witness deque: the reference is unsatisfiable (panics) here and
             * nothing meaningful follows; modelled as: report it, the FSM idles from this cycle on */
            fail(&st, (int64_t)cyc, ZKC_CU_CHK_WITNESS_EXHAUSTED);
            s.state_get_from_queue = 0;
        }
        if (s.state_get_from_queue) { /* :195-196 */
            req = requests[rpos++]; req._pad = 0; req.is_first &= 1;
            uint64_t enc[8];
            orc_decommit_query_encode(&req, enc);
            memcpy(rq.head, enc, 64);
            orc_poseidon2_permutation(rq.head);
            rq.length--;
        }
        const uint32_t top = req.code_hash[7];
        const int version_matches = (top >> 16) == ZKC_CODE_HASH_VERSION_TOP16;
        if (s.state_get_from_queue && !version_matches) fail(&st, (int64_t)cyc, ZKC_CU_CHK_VERSION); /* :202-204 */
        const uint32_t length_in_words = s.state_get_from_queue ? (top & 0xFFFF) : 1;
        if ((length_in_words + 1) & 1) fail(&st, (int64_t)cyc, ZKC_CU_CHK_LENGTH); /* :215-221: the halved value must be a UInt16 */
        const uint32_t length_in_rounds = (length_in_words + 1) >> 1;
        if (s.state_get_from_queue) { /* :233-275 */
            s.num_rounds_left = length_in_rounds;
            s.length_in_bits = length_in_words * 256;
            s.timestamp = req.timestamp;
            s.current_page = req.page;
            memcpy(s.hash_to_compare_against, req.code_hash, 28);
            s.hash_to_compare_against[7] = 0;
            s.current_index = 0;
            memcpy(s.sha256_inner_state, ORC_SHA256_IV, 32);
        }
        s.state_decommit = s.state_decommit || s.state_get_from_queue;
        s.state_get_from_queue = 0;
        if (s.state_decommit) s.num_rounds_left = (s.num_rounds_left - 1) & 0xFFFF; /* :281-287 */
        const int last_round = s.num_rounds_left == 0;
        const int finalize = last_round && s.state_decommit;
        const int process_second_word = !last_round && s.state_decommit;
        uint32_t w0[8] = {0}, w1[8] = {0};
        if (s.state_decommit) { /* :295-304 */
            if (wpos < n_code_words) memcpy(w0, code_words + 8 * wpos, 32);
            else fail(&st, (int64_t)cyc, ZKC_CU_CHK_WITNESS_EXHAUSTED);
            wpos++;
        }
        if (process_second_word) {
            if (wpos < n_code_words) memcpy(w1, code_words + 8 * wpos, 32);
            else fail(&st, (int64_t)cyc, ZKC_CU_CHK_WITNESS_EXHAUSTED);
            wpos++;
        }
        const uint32_t index0 = s.current_index;
        if (s.state_decommit) s.current_index++;
        const uint32_t index1 = s.current_index;
        if (process_second_word) s.current_index++;
        memory_push(&mq, s.timestamp, s.current_page, index0, w0, s.state_decommit, memory_states, &n_states);
        zkc_queue_state12 mq_after0 = mq;
        memory_push(&mq, s.timestamp, s.current_page, index1, w1, process_second_word, memory_states, &n_states);
        uint32_t m[16]; /* :354-381, big-endian words: limb 7 first */
        for (int i = 0; i < 8; i++) { m[i] = w0[7 - i]; m[8 + i] = w1[7 - i]; }
        if (finalize) {
            m[8] = 0x80000000u;
            for (int i = 9; i < 15; i++) m[i] = 0;
            m[15] = s.length_in_bits;
        }
        uint32_t state_in[8], ns[8];
        memcpy(state_in, s.sha256_inner_state, 32);
        memcpy(ns, s.sha256_inner_state, 32);
        orc_sha256_compress(ns, m);
        if (s.state_decommit) memcpy(s.sha256_inner_state, ns, 32);
        if (finalize) /* :393-420: hash = [ns7 .. ns1, 0] as little-endian limbs */
            for (int i = 0; i < 7; i++)
                if (ns[7 - i] != s.hash_to_compare_against[i]) fail(&st, (int64_t)cyc, ZKC_CU_CHK_HASH);
        if (finalize && s.hash_to_compare_against[7] != 0) fail(&st, (int64_t)cyc, ZKC_CU_CHK_HASH);
        const int is_empty = rq.length == 0;
        s.finished = s.finished || (is_empty && finalize);
        s.state_get_from_queue = !is_empty && finalize;
        s.state_decommit = process_second_word;

        if (trace) {
            T(ZKC_CU_FLAGS_IN + 0, cyc) = in.state_get_from_queue; T(ZKC_CU_FLAGS_IN + 1, cyc) = in.state_decommit;
            T(ZKC_CU_FLAGS_IN + 2, cyc) = in.finished;
            uint64_t flat[11];
            orc_decommit_query_flatten(&req, flat);
            for (int i = 0; i < 11; i++) T(ZKC_CU_REQUEST + i, cyc) = flat[i];
            for (int i = 0; i < 12; i++) T(ZKC_CU_REQ_HEAD + i, cyc) = rq.head[i];
            T(ZKC_CU_REQ_LEN, cyc) = rq.length;
            T(ZKC_CU_VERSION_MATCHES, cyc) = (uint64_t)version_matches;
            T(ZKC_CU_LENGTH_IN_WORDS, cyc) = length_in_words; T(ZKC_CU_LENGTH_IN_ROUNDS, cyc) = length_in_rounds;
            T(ZKC_CU_LENGTH_IN_BITS, cyc) = s.length_in_bits; T(ZKC_CU_TIMESTAMP, cyc) = s.timestamp; T(ZKC_CU_PAGE, cyc) = s.current_page;
            for (int i = 0; i < 8; i++) T(ZKC_CU_HASH_TO_COMPARE + i, cyc) = s.hash_to_compare_against[i];
            T(ZKC_CU_DECOMMIT, cyc) = (uint64_t)(finalize || process_second_word);
            T(ZKC_CU_NUM_ROUNDS_LEFT, cyc) = s.num_rounds_left; T(ZKC_CU_LAST_ROUND, cyc) = (uint64_t)last_round;
            T(ZKC_CU_FINALIZE, cyc) = (uint64_t)finalize; T(ZKC_CU_PROCESS_SECOND_WORD, cyc) = (uint64_t)process_second_word;
            for (int i = 0; i < 8; i++) { T(ZKC_CU_WORD0 + i, cyc) = w0[i]; T(ZKC_CU_WORD1 + i, cyc) = w1[i]; }
            T(ZKC_CU_INDEX0, cyc) = index0; T(ZKC_CU_INDEX1, cyc) = index1; T(ZKC_CU_INDEX_OUT, cyc) = s.current_index;
            for (int i = 0; i < 12; i++) { T(ZKC_CU_MEM_TAIL0 + i, cyc) = mq_after0.tail[i]; T(ZKC_CU_MEM_TAIL1 + i, cyc) = mq.tail[i]; }
            T(ZKC_CU_MEM_TAIL0 + 12, cyc) = mq_after0.length; T(ZKC_CU_MEM_TAIL1 + 12, cyc) = mq.length;
            for (int i = 0; i < 16; i++) T(ZKC_CU_MESSAGE + i, cyc) = m[i];
            for (int i = 0; i < 8; i++) {
                T(ZKC_CU_STATE_IN + i, cyc) = state_in[i]; T(ZKC_CU_STATE_NEW + i, cyc) = ns[i];
                T(ZKC_CU_STATE_OUT + i, cyc) = s.sha256_inner_state[i];
            }
            T(ZKC_CU_FLAGS_OUT + 0, cyc) = s.state_get_from_queue; T(ZKC_CU_FLAGS_OUT + 1, cyc) = s.state_decommit;
            T(ZKC_CU_FLAGS_OUT + 2, cyc) = s.finished;
        }
    }
    if (n_memory_states) *n_memory_states = n_states;
    /* :449 enforce_consistency */
    if (rq.length == 0 && memcmp(rq.head, rq.tail, 96)) fail(&st, -1, ZKC_CU_CHK_QUEUE_CONSISTENCY);
    const int done = s.finished != 0; /* :113-115 */

    zkc_code_unpacker_fsm out;
    memset(&out, 0, sizeof out);
    out.internal_fsm = s;
    out.decommittment_requests_queue_state = rq;
    out.memory_queue_state = mq;
    zkc_queue_state12 obs_out;
    memset(&obs_out, 0, sizeof obs_out);
    if (done) obs_out = mq; /* :118-123 */

    uint64_t e_in[50], e_out[25], e_fin[74], e_fout[74];
    size_t n_in = put_queue_state12(e_in, &io->memory_queue_initial_state);
    n_in += put_queue_state12(e_in + n_in, &io->sorted_requests_queue_initial_state);
    const size_t n_out = put_queue_state12(e_out, &obs_out);
    const size_t n_fin = orc_code_unpacker_encode_fsm(fin, e_fin);
    const size_t n_fout = orc_code_unpacker_encode_fsm(&out, e_fout);
    if (options && options->compare_expected) {
        uint64_t b[74], c[25];
        orc_code_unpacker_encode_fsm(&io->hidden_fsm_output, b);
        put_queue_state12(c, &io->memory_queue_final_state);
        if (memcmp(e_fout, b, sizeof b) || memcmp(e_out, c, sizeof c) || (io->completion_flag != 0) != done)
            if (st.code == ZKC_OK) st.code = ZKC_ERR_FSM_OUTPUT_MISMATCH;
    }
    io->hidden_fsm_output = out;
    io->memory_queue_final_state = obs_out;
    io->completion_flag = (uint32_t)done;
    orc_closed_form_commitment(start, done, e_in, n_in, e_out, n_out, e_fin, n_fin, e_fout, n_fout, commitment);
    if (status) *status = st;
    return st.code;
}
