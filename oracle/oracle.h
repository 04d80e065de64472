/* ORACLE -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, sequential, one loop iteration at a time exactly like the reference)
 * of the hot path of matter-labs/era-zkevm_circuits.  It is the parity checker for the CUDA engine
 * and the `cpu_baseline` of bench.py.  The product (era_zkevm_circuits_b200/) never links, imports
 * or executes anything in this directory.  It shares only the record/column declarations of the
 * public ABI header (include/zkc_b200.h) so that both sides speak about the same cells.
 *
 * Pinning status (details in each file and DESIGN.md):
 *   Goldilocks field         -- pinned by definition (Python big-int cross-check)
 *   keccak-f / Keccak-256    -- pinned (reference tests compare against sha3::Keccak256)
 *   SHA-256 compression      -- pinned against hashlib
 *   Poseidon2                -- PARITY UNPINNED (un-vendored boojum; no reference KAT)
 *   sorter / RAM loop logic  -- pinned to "the reference's own test vectors satisfy every
 *                               enforcement"; accumulator values are unpinned (depend on Poseidon2)
 */
#ifndef ORC_ORACLE_H
#define ORC_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#include "../include/zkc_b200.h"
#include "gl.h"

#define ORC_P2_NUM_CONSTANTS 360

void orc_poseidon2_constants(uint64_t out[ORC_P2_NUM_CONSTANTS]);
void orc_poseidon2_permutation(uint64_t s[12]);
void orc_sponge_init(uint64_t s[12], uint64_t length);
void orc_sponge_absorb8(uint64_t s[12], const uint64_t chunk[8]);
void orc_commit_encoding(const uint64_t *input, size_t n, uint64_t out[4]);
void orc_closed_form_commitment(int start_flag, int completion_flag, const uint64_t *obs_in, size_t n_obs_in,
                                const uint64_t *obs_out, size_t n_obs_out, const uint64_t *fsm_in,
                                size_t n_fsm_in, const uint64_t *fsm_out, size_t n_fsm_out, uint64_t out[4]);
void orc_produce_fs_challenges(const uint64_t *unsorted_tail, uint32_t unsorted_len, const uint64_t *sorted_tail,
                               uint32_t sorted_len, int tw, int num_challenges, uint64_t *result);

/* field helpers exported for the Python tests */
uint64_t orc_gl_mul(uint64_t a, uint64_t b);
uint64_t orc_gl_add(uint64_t a, uint64_t b);
uint64_t orc_gl_sub(uint64_t a, uint64_t b);
uint64_t orc_gl_inv(uint64_t a);

/* encodings */
void orc_memory_query_encode(const zkc_memory_query *q, uint64_t out[8]);

/* utils.rs:81-137 on column-major inputs; same contract as zkc_accumulate_grand_products */
void orc_accumulate_grand_products(const uint64_t *lhs_enc, const uint64_t *rhs_enc, const uint8_t *should_acc,
                                   size_t enc_len, size_t rows, const uint64_t *challenges,
                                   const uint64_t acc_in[4], uint64_t *acc_out, uint64_t *chain_out,
                                   uint64_t acc_final[4]);

/* pushes `n` queries into an empty full-state queue (FullStateCircuitQueue::push), recording the
 * tail BEFORE each push into prev_states[n][12] (the raw witness' second tuple element) and the
 * final state. Used to build inputs exactly like the reference test does (:506-515). */
void orc_memory_queue_simulate(const zkc_memory_query *q, size_t n, uint64_t *prev_states, zkc_queue_state12 *final_state);

size_t orc_ram_encode_input_data(const zkc_ram_input_data *d, uint64_t *dst);
size_t orc_ram_encode_fsm(const zkc_ram_fsm *f, uint64_t *dst);
int orc_ram_permutation_entry_point(zkc_ram_closed_form *io, const zkc_memory_query *unsorted, size_t n_unsorted,
                                    const zkc_memory_query *sorted, size_t n_sorted, size_t limit,
                                    const zkc_ram_options *options, uint64_t *trace, uint64_t commitment[4],
                                    zkc_status *status);


/* log_query.c */
void orc_log_query_encode(const zkc_log_query *q, uint64_t out[20]);
void orc_log_query_flatten(const zkc_log_query *q, uint64_t out[36]);
void orc_log_queue_absorb(uint64_t chain[4], const uint64_t enc[20], uint64_t *rounds);
void orc_log_queue_simulate(const zkc_log_query *q, const uint32_t *extra_ts, size_t n, uint64_t *prev_tails,
                            zkc_queue_state4 *final_state);
size_t orc_put_queue_state4(uint64_t *dst, const zkc_queue_state4 *s);
/* log_sorter.c; result_tails (optional out): tail after each executed push, n_result_tails their count */
size_t orc_events_encode_fsm(const zkc_events_fsm *f, uint64_t *dst);
int orc_log_sorter_entry_point(zkc_events_closed_form *io, const zkc_log_query *unsorted, size_t n_unsorted,
                               const zkc_log_query *sorted, size_t n_sorted, size_t limit,
                               const zkc_sorter_options *options, uint64_t *trace, uint64_t *result_tails,
                               size_t *n_result_tails, uint64_t commitment[4], zkc_status *status);
/* storage_validity.c */
size_t orc_storage_encode_fsm(const zkc_storage_fsm *f, uint64_t *dst);
int orc_storage_validity_entry_point(zkc_storage_closed_form *io, const zkc_log_query *unsorted, size_t n_unsorted,
                                     const zkc_log_query *sorted, const uint32_t *sorted_ts, size_t n_sorted, size_t limit,
                                     const zkc_sorter_options *options, uint64_t *trace, uint64_t *result_tails,
                                     size_t *n_result_tails, uint64_t commitment[4], zkc_status *status);
/* sort_decommittments.c; result_states (optional out): result-queue tail [12] after each executed push */
void orc_decommit_query_encode(const zkc_decommit_query *q, uint64_t out[8]);
void orc_decommit_query_flatten(const zkc_decommit_query *q, uint64_t out[11]);
void orc_decommit_queue_simulate(const zkc_decommit_query *q, size_t n, uint64_t *prev_states, zkc_queue_state12 *final_state);
size_t orc_decommit_sorter_encode_fsm(const zkc_decommit_sorter_fsm *f, uint64_t *dst);
int orc_sort_decommittments_entry_point(zkc_decommit_sorter_closed_form *io, const zkc_decommit_query *unsorted, size_t n_unsorted,
                                        const zkc_decommit_query *sorted, size_t n_sorted, size_t limit,
                                        const zkc_sorter_options *options, uint64_t *trace, uint64_t *result_states,
                                        size_t *n_result_states, uint64_t commitment[4], zkc_status *status);
/* demux_log_queue.c; output_tails (optional out): [6][limit][4], queue q's tail after each of its executed pushes */
size_t orc_demux_encode_fsm(const zkc_demux_fsm *f, uint64_t *dst);
int orc_demux_log_queue_entry_point(zkc_demux_closed_form *io, const zkc_log_query *records, size_t n_records, size_t limit,
                                    const zkc_demux_options *options, uint64_t *trace, uint64_t *output_tails,
                                    size_t n_output_tails[6], uint64_t commitment[4], zkc_status *status);
/* linear_hasher.c; keccak_states (optional out): [limit][25], the keccak state after every cycle */
int orc_log_query_into_bytes(const zkc_log_query *q, uint8_t out[ZKC_LH_MESSAGE_BYTES]);
int orc_linear_hasher_entry_point(zkc_linear_hasher_closed_form *io, const zkc_log_query *records, size_t n_records, size_t limit,
                                  const zkc_sorter_options *options, uint64_t *trace, uint64_t *keccak_states, uint64_t commitment[4],
                                  zkc_status *status);
/* code_unpacker_sha256.c; memory_states (optional out): memory queue tail [12] after each executed push */
size_t orc_code_unpacker_encode_fsm(const zkc_code_unpacker_fsm *f, uint64_t *dst);
int orc_code_unpacker_entry_point(zkc_code_unpacker_closed_form *io, const zkc_decommit_query *requests, size_t n_requests,
                                  const uint32_t *code_words, size_t n_code_words, size_t limit, const zkc_sorter_options *options,
                                  uint64_t *trace, uint64_t *memory_states, size_t *n_memory_states, uint64_t commitment[4],
                                  zkc_status *status);
/* keccak256_round_function.c; memory_states (optional out): memory queue tail after each executed push */
void orc_keccak_f1600(uint64_t A[25]);
void orc_keccak256(const uint8_t *msg, size_t len, uint8_t digest[32]);
size_t orc_keccak_encode_fsm(const zkc_keccak_fsm *f, uint64_t *dst);
int orc_keccak256_entry_point(zkc_keccak_closed_form *io, const zkc_log_query *requests, size_t n_requests,
                              const uint32_t *memory_reads, size_t n_reads, size_t limit,
                              const zkc_precompile_options *options, uint64_t *trace, uint64_t *memory_states,
                              size_t *n_memory_states, uint64_t commitment[4], zkc_status *status);
/* sha256_round_function.c */
void orc_sha256_compress(uint32_t state[8], const uint32_t m[16]);
size_t orc_sha256_encode_fsm(const zkc_sha256_fsm *f, uint64_t *dst);
int orc_sha256_entry_point(zkc_sha256_closed_form *io, const zkc_log_query *requests, size_t n_requests,
                           const uint32_t *memory_reads, size_t n_reads, size_t limit, const zkc_precompile_options *options,
                           uint64_t *trace, uint64_t *memory_states, size_t *n_memory_states, uint64_t commitment[4],
                           zkc_status *status);
/* main_vm.c */
size_t orc_vm_flatten_state(const zkc_vm_state *s, uint64_t *dst);
void orc_vm_context_encode(const zkc_vm_context *c, uint64_t e[32]);
void orc_vm_initial_bootloader_state(const zkc_vm_closed_form *io, const zkc_vm_isa *isa, zkc_vm_state *st);
int orc_main_vm_run(const zkc_vm_isa *isa, const zkc_vm_closed_form *gc, const zkc_vm_state *initial, const uint32_t *code, size_t code_words,
                    size_t cycles, zkc_vm_state *snapshots, zkc_vm_cycle_witness *witness, zkc_vm_callstack_witness *cw_out,
                    size_t cw_cap, size_t *n_cw, uint64_t rollback_tail_out[4], zkc_status *status);
int orc_main_vm_entry_point(zkc_vm_closed_form *io, const zkc_vm_isa *isa, const zkc_vm_state *snapshots,
                            const zkc_vm_cycle_witness *witness, const zkc_vm_callstack_witness *callstack_witness,
                            size_t n_callstack_witness, size_t limit, const zkc_vm_options *options,
                            uint64_t *trace, uint64_t commitment[4], zkc_status *status);
/* main_vm_gadgets.c: trace [n_instances][ZKC_VM_NUM_COLS][limit] -> out [n_instances][ZKC_VMG_NUM_COLS][limit] */
void orc_main_vm_gadget_cells(const uint64_t *trace, size_t limit, size_t n_instances, uint64_t *out);
/* the ptr / jump / context block: + snapshots [n_instances][limit + 1] -> out [n_instances][ZKC_VMS_NUM_COLS][limit] */
void orc_main_vm_state_gadget_cells(const uint64_t *trace, const zkc_vm_state *snapshots, size_t limit, size_t n_instances, uint64_t *out);
/* the fetch / src0 read / dst0 write memory-queue relations of every cycle -> out [n_instances][ZKC_VMQ_NUM_COLS][limit] */
void orc_main_vm_memory_sponge_cells(const uint64_t *trace, const zkc_vm_state *snapshots, size_t limit, size_t n_instances, uint64_t *out);
/* the cells of create_prestate that are not DENSE columns -> out [n_instances][ZKC_VMP_NUM_COLS][limit] */
void orc_main_vm_prestate_cells(const uint64_t *trace, const zkc_vm_state *snapshots, size_t limit, size_t n_instances, uint64_t *out);
/* the register write-back of the state diffs (cycle.rs:158-433) -> out [n_instances][ZKC_VMW_NUM_COLS][limit] */
void orc_main_vm_writeback_cells(const zkc_vm_isa *isa, const uint64_t *trace, const zkc_vm_state *snapshots, size_t limit, size_t n_instances,
                                 uint64_t *out);
#endif
