/* ORACLE -- TEST INFRASTRUCTURE ONLY (see gl.h header).
 *
 * Sequential CPU restatement of the log-queue demultiplexer:
 *   demultiplex_storage_logs_enty_point   /root/reference/src/demux_log_queue/mod.rs:38-217
 *   demultiplex_storage_logs_inner        /root/reference/src/demux_log_queue/mod.rs:234-399
 *   push_with_optimize                    /root/reference/src/demux_log_queue/mod.rs:401-444
 *   check_if_bitmask_and_if_empty         /root/reference/src/demux_log_queue/mod.rs:446-459
 * The constants compared with (aux bytes, formal precompile addresses) live in the un-vendored zkevm_opcode_defs
 * (system_params, v1.4.1): STORAGE/EVENT/L1_MESSAGE/PRECOMPILE_AUX_BYTE = 0/1/2/3, keccak256 0x8010, sha256 0x02,
 * ecrecover 0x01.
 * Pinning: loop logic pinned by the reference's test vector (mod.rs:602-923, limit 16: every enforcement holds);
 * hash-dependent values PARITY UNPINNED (Poseidon2, see poseidon2.c).
 */
#include "oracle.h"
#include <string.h>

static void fail(zkc_status *st, int64_t row, uint32_t bit) {
    st->code = ZKC_ERR_UNSATISFIED;
    st->failed_checks |= bit;
    if (row >= 0 && (st->first_bad_row < 0 || row < st->first_bad_row)) st->first_bad_row = row;
}

/* CSVarLengthEncodable order of LogDemuxerFSMInputOutput, input.rs:24-32 */
size_t orc_demux_encode_fsm(const zkc_demux_fsm *f, uint64_t *dst) {
    size_t n = orc_put_queue_state4(dst, &f->initial_log_queue_state);
    for (int q = 0; q < ZKC_DEMUX_NUM_QUEUES; q++) n += orc_put_queue_state4(dst + n, &f->output_queue_states[q]);
    return n; /* 63 */
}

#define T(col, r) trace[(size_t)(col) * limit + (r)]

/* output_tails (optional out): [6][limit][4], queue q's tail after each of its executed pushes; n_output_tails[6] */
int orc_demux_log_queue_entry_point(zkc_demux_closed_form *io, const zkc_log_query *records, size_t n_records, size_t limit,
                                    const zkc_demux_options *options, uint64_t *trace, uint64_t *output_tails,
                                    size_t n_output_tails[6], uint64_t commitment[4], zkc_status *status) {
    zkc_status st = {ZKC_OK, 0, -1, 0, 0};
    uint32_t aux_bytes[4] = {0, 1, 2, 3}, addresses[3] = {0x8010, 0x02, 0x01};
    if (options && options->custom_constants) {
        memcpy(aux_bytes, options->aux_bytes, sizeof aux_bytes);
        memcpy(addresses, options->precompile_addresses, sizeof addresses);
    }
    const int start = io->start_flag != 0;
    const zkc_demux_fsm *fin = &io->hidden_fsm_input;
    static const uint64_t zero4[4] = {0, 0, 0, 0};
    if (memcmp(io->initial_log_queue_state.head, zero4, 32)) fail(&st, -1, ZKC_DMX_CHK_TRIVIAL_HEAD); /* :66-69 */
    zkc_queue_state4 iq = start ? io->initial_log_queue_state : fin->initial_log_queue_state;
    zkc_queue_state4 oq[6];
    memset(oq, 0, sizeof oq);
    if (!start) memcpy(oq, fin->output_queue_states, sizeof oq); /* :83-106 */
    size_t pushes[6] = {0, 0, 0, 0, 0, 0};

    size_t pos = 0;
    for (size_t cyc = 0; cyc < limit; cyc++) {
        const int queue_is_empty = iq.length == 0, execute = !queue_is_empty;
        zkc_log_query it;
        memset(&it, 0, sizeof it);
        if (execute && pos < n_records) it = records[pos++];
        uint64_t enc[20], rounds[36];
        orc_log_query_encode(&it, enc);
        if (execute) { orc_log_queue_absorb(iq.head, enc, NULL); iq.length--; }
        const uint32_t aux = ZKC_LQ_AUX(it.flags);
        int is_aux[4], is_addr[3];
        for (int i = 0; i < 4; i++) is_aux[i] = aux == aux_bytes[i];
        for (int i = 0; i < 3; i++)
            is_addr[i] = it.address[0] == addresses[i] && !it.address[1] && !it.address[2] && !it.address[3] && !it.address[4];
        const int is_rollup_shard = ZKC_LQ_SHARD(it.flags) == 0;
        const int execute_porter_storage = is_aux[0] && !is_rollup_shard && execute;
        if (execute_porter_storage) fail(&st, (int64_t)cyc, ZKC_DMX_CHK_PORTER_STORAGE); /* :304-305 */
        const int bitmask[6] = {is_aux[0] && is_rollup_shard && execute, is_aux[1] && execute, is_aux[2] && execute,
                                is_aux[3] && is_addr[0] && execute, is_aux[3] && is_addr[1] && execute,
                                is_aux[3] && is_addr[2] && execute};
        /* push_with_optimize: the state of the last queue whose bit is set, else of queue 0 */
        int sel = 0;
        for (int q = 1; q < 6; q++) if (bitmask[q]) sel = q;
        const zkc_queue_state4 exec_before = oq[sel];
        uint64_t exec_tail[4];
        memcpy(exec_tail, exec_before.tail, 32);
        orc_log_queue_absorb(exec_tail, enc, rounds);
        const uint32_t exec_len = exec_before.length + 1;
        for (int q = 0; q < 6; q++)
            if (bitmask[q]) {
                memcpy(oq[q].tail, exec_tail, 32);
                oq[q].length = exec_len;
                if (output_tails) memcpy(output_tails + 4 * ((size_t)q * limit + pushes[q]), exec_tail, 32);
                pushes[q]++;
            }
        const int is_bitmask = is_aux[0] + is_aux[1] + is_aux[2] + is_aux[3] == 1;
        if (execute && !is_bitmask) fail(&st, (int64_t)cyc, ZKC_DMX_CHK_BITMASK); /* :383-384 */

        if (trace) {
            T(ZKC_DMX_QUEUE_IS_EMPTY, cyc) = (uint64_t)queue_is_empty; T(ZKC_DMX_EXECUTE, cyc) = (uint64_t)execute;
            uint64_t flat[36];
            orc_log_query_flatten(&it, flat);
            for (int i = 0; i < 36; i++) T(ZKC_DMX_ITEM + i, cyc) = flat[i];
            for (int i = 0; i < 20; i++) T(ZKC_DMX_ENC + i, cyc) = enc[i];
            for (int i = 0; i < 4; i++) T(ZKC_DMX_HEAD + i, cyc) = iq.head[i];
            T(ZKC_DMX_LEN, cyc) = iq.length;
            for (int i = 0; i < 4; i++) T(ZKC_DMX_IS_AUX + i, cyc) = (uint64_t)is_aux[i];
            for (int i = 0; i < 3; i++) T(ZKC_DMX_IS_ADDRESS + i, cyc) = (uint64_t)is_addr[i];
            T(ZKC_DMX_IS_ROLLUP_SHARD, cyc) = (uint64_t)is_rollup_shard;
            T(ZKC_DMX_EXECUTE_PORTER_STORAGE, cyc) = (uint64_t)execute_porter_storage;
            for (int q = 0; q < 6; q++) T(ZKC_DMX_BITMASK + q, cyc) = (uint64_t)bitmask[q];
            T(ZKC_DMX_IS_BITMASK, cyc) = (uint64_t)is_bitmask;
            for (int i = 0; i < 4; i++) T(ZKC_DMX_EXEC_TAIL + i, cyc) = exec_before.tail[i];
            T(ZKC_DMX_EXEC_LEN, cyc) = exec_before.length;
            for (int i = 0; i < 36; i++) T(ZKC_DMX_PUSH_ROUND0 + i, cyc) = rounds[i];
            for (int q = 0; q < 6; q++) {
                for (int i = 0; i < 4; i++) T(ZKC_DMX_QUEUE_TAILS + 4 * q + i, cyc) = oq[q].tail[i];
                T(ZKC_DMX_QUEUE_LENS + q, cyc) = oq[q].length;
            }
        }
    }
    if (n_output_tails) for (int q = 0; q < 6; q++) n_output_tails[q] = pushes[q];
    /* :395 enforce_consistency */
    if (iq.length == 0 && memcmp(iq.head, iq.tail, 32)) fail(&st, -1, ZKC_DMX_CHK_QUEUE_CONSISTENCY);
    const int completed = iq.length == 0; /* :118-119 */

    zkc_demux_fsm out;
    memset(&out, 0, sizeof out);
    out.initial_log_queue_state = iq;
    memcpy(out.output_queue_states, oq, sizeof oq);
    zkc_queue_state4 obs_out[6];
    memset(obs_out, 0, sizeof obs_out);
    if (completed) memcpy(obs_out, oq, sizeof oq); /* :144-199 */

    uint64_t e_in[9], e_out[54], e_fin[63], e_fout[63];
    const size_t n_in = orc_put_queue_state4(e_in, &io->initial_log_queue_state);
    size_t n_out = 0;
    for (int q = 0; q < 6; q++) n_out += orc_put_queue_state4(e_out + n_out, &obs_out[q]);
    const size_t n_fin = orc_demux_encode_fsm(fin, e_fin);
    const size_t n_fout = orc_demux_encode_fsm(&out, e_fout);
    if (options && options->compare_expected) {
        uint64_t b[63], c[54];
        orc_demux_encode_fsm(&io->hidden_fsm_output, b);
        size_t n = 0;
        for (int q = 0; q < 6; q++) n += orc_put_queue_state4(c + n, &io->output_queue_states[q]);
        if (memcmp(e_fout, b, sizeof b) || memcmp(e_out, c, sizeof c) || (io->completion_flag != 0) != completed)
            if (st.code == ZKC_OK) st.code = ZKC_ERR_FSM_OUTPUT_MISMATCH;
    }
    io->hidden_fsm_output = out;
    memcpy(io->output_queue_states, obs_out, sizeof obs_out);
    io->completion_flag = (uint32_t)completed;
    orc_closed_form_commitment(start, completed, e_in, n_in, e_out, n_out, e_fin, n_fin, e_fout, n_fout, commitment);
    if (status) *status = st;
    return st.code;
}
