/* zkc_b200 -- C ABI of the B200-native witness-generation / constraint-evaluation engine for the
 * zkSync Era zkEVM circuits' data-parallel hot path.
 *
 * The reference (matter-labs/era-zkevm_circuits, /root/reference) has no FFI of its own: its only
 * seam is the generic `*_entry_point(cs, witness, round_function, limit) -> [Num<F>; 4]` function
 * per circuit.  Each `zkc_*_entry_point` below replaces the body of the reference function cited
 * next to it: the Rust shim marshals the `*CircuitInstanceWitness` into the
 * `#[repr(C)]` records declared here, makes ONE call, and bulk-assigns the returned witness
 * columns to its boojum variables.  Plain pointers and sizes only; no callbacks into the host
 * during a call; failures are status codes (the reference panics / becomes unsatisfiable).
 *
 * Conventions
 *  - field elements are canonical Goldilocks values (< p = 2^64 - 2^32 + 1) in uint64_t;
 *  - "trace" outputs are COLUMN-MAJOR: cell (col, row) lives at trace[col * limit + row], one row per
 *    iteration of the reference's `for _cycle in 0..limit` loop, one column per value the reference
 *    names in that loop body, in the order the loop body allocates them;
 *  - `on_device != 0` means every bulk pointer argument (records, prev states, trace) is a device
 *    pointer already resident in HBM; otherwise they are host pointers and the call does the
 *    H2D/D2H copies itself (pinned host memory makes them asynchronous).  The circuit entry points
 *    take it as a bit mask: ZKC_INPUTS_ON_DEVICE | ZKC_TRACE_ON_DEVICE (e.g. host inputs, witness
 *    left in HBM for a GPU prover = ZKC_TRACE_ON_DEVICE);
 *  - small structs (closed-form inputs/outputs, status) are always host memory.
 */
#ifndef ZKC_B200_H
#define ZKC_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZKC_GL_P 0xFFFFFFFF00000001ULL
#define ZKC_INPUTS_ON_DEVICE 1
#define ZKC_TRACE_ON_DEVICE 2
#define ZKC_ALL_ON_DEVICE 3
#define ZKC_NUM_REPETITIONS 2 /* DEFAULT_NUM_PERMUTATION_ARGUMENT_REPETITIONS, src/lib.rs:39 */
#define ZKC_COMMITMENT_LEN 4  /* INPUT_OUTPUT_COMMITMENT_LENGTH, src/fsm_input_output/circuit_inputs/mod.rs:4 */
#define ZKC_FULL_STATE 12     /* FULL_SPONGE_QUEUE_STATE_WIDTH, src/base_structures/vm_state/mod.rs:27-30 */
#define ZKC_QUEUE_STATE 4     /* QUEUE_STATE_WIDTH */
#define ZKC_BOOTLOADER_HEAP_PAGE_DEFAULT 10u /* zkevm_opcode_defs::BOOTLOADER_HEAP_PAGE (un-vendored; from memory) */

/* ---- status ------------------------------------------------------------------------------ */
enum zkc_code {
    ZKC_OK = 0,
    ZKC_ERR_INVALID_ARGUMENT = 1,
    ZKC_ERR_CUDA = 2,
    ZKC_ERR_NO_DEVICE = 3,
    /* the reference would produce an unsatisfiable constraint system (an `enforce_*` fails) */
    ZKC_ERR_UNSATISFIED = 4,
    /* the reference would panic in hook_compare_witness (fsm_input_output/mod.rs:102-133) */
    ZKC_ERR_FSM_OUTPUT_MISMATCH = 5,
    /* raw queue witness is not a hash chain (previous-state column inconsistent) */
    ZKC_ERR_QUEUE_WITNESS_INCONSISTENT = 6,
};

/* which enforcement failed first (row-ordered); bit set per failing check, circuit-specific */
typedef struct zkc_status {
    int32_t code;           /* enum zkc_code */
    int32_t cuda_error;     /* cudaError_t when code == ZKC_ERR_CUDA */
    int64_t first_bad_row;  /* loop iteration of the first failing enforcement, -1 if none/global */
    uint32_t failed_checks; /* bit mask, see ZKC_*_CHK_* */
    uint32_t reserved;
} zkc_status;

typedef struct zkc_ctx zkc_ctx;

/* create an engine bound to one GPU (one process per GPU); uploads round constants */
int zkc_create(int device, zkc_ctx **out);
void zkc_destroy(zkc_ctx *ctx);
/* all work of later calls is enqueued on `cuda_stream` (a cudaStream_t; NULL = legacy default) */
int zkc_set_stream(zkc_ctx *ctx, void *cuda_stream);
const char *zkc_version(void);
/* kernels launched by this context since creation (the bench's `gpu_launches` claim) */
uint64_t zkc_launch_count(const zkc_ctx *ctx);
/* streaming multiprocessors of the bound device (grid sizing of callers that batch instances) */
int zkc_sm_count(const zkc_ctx *ctx);
/* per-kernel device timing (CUDA events on the context's stream). enable!=0 starts recording. */
int zkc_profile_enable(zkc_ctx *ctx, int enable);
/* sums elapsed ms / launches recorded for kernel `name` since the last reset; resolves events */
int zkc_profile_query(zkc_ctx *ctx, const char *name, double *ms_total, uint64_t *launches);
int zkc_profile_reset(zkc_ctx *ctx);
/* pinned host allocations for asynchronous copies (cudaHostAlloc / cudaFreeHost) */
void *zkc_host_alloc(size_t bytes);
void zkc_host_free(void *p);

/* ---- primitives (device buffers or host buffers per on_device) --------------------------------- */
/* Poseidon2 permutation of n independent 12-element states, AoS [n][12].
 * Replaces R::compute_round_function (boojum, called at src/main_vm/utils.rs:212, cycle.rs:952). */
int zkc_poseidon2_permute(zkc_ctx *ctx, const uint64_t *states_in, uint64_t *states_out, size_t n,
                          int on_device);
/* commit_encoding, src/fsm_input_output/mod.rs:281-326, for `n_items` inputs of `len` elements
 * each (AoS [n_items][len]); out AoS [n_items][4]. */
int zkc_commit_encoding(zkc_ctx *ctx, const uint64_t *inputs, size_t len, size_t n_items,
                        uint64_t *out, int on_device);

/* element-wise Goldilocks a*b, a+b, a-b, a*b+c on canonical inputs (host buffers).  Diagnostic entry:
 * lets the tests pin the register-level field arithmetic (Num::{mul,add,sub,fma} of boojum, call sites
 * src/utils.rs:112-132) against big-integer arithmetic, edge values included. */
int zkc_field_ops(zkc_ctx *ctx, const uint64_t *a, const uint64_t *b, const uint64_t *c, size_t n,
                  uint64_t *out_mul, uint64_t *out_add, uint64_t *out_sub, uint64_t *out_fma);

/* accumulate_grand_products<ENC, ENC+1, 2>, src/utils.rs:81-137, over `rows` loop iterations.
 *   lhs_enc/rhs_enc : column-major [enc_len][rows]
 *   should_acc      : [rows] of 0/1 (NULL = all ones)
 *   challenges      : host, [2][enc_len + 1] as returned by produce_fs_challenges
 *   acc_in          : host, lhs[2] then rhs[2] (initial accumulators)
 *   acc_out         : column-major [4][rows] = lhs rep0, lhs rep1, rhs rep0, rhs rep1 after each row
 *   chain_out       : NULL, or column-major [4 * enc_len][rows]: the Num::fma partial sums
 *                     (utils.rs:112-128), column (rep*2 + side) * enc_len + i, side 0 = lhs
 *   acc_final       : host, the 4 accumulators after the last row */
int zkc_accumulate_grand_products(zkc_ctx *ctx, const uint64_t *lhs_enc, const uint64_t *rhs_enc,
                                  const uint8_t *should_acc, size_t enc_len, size_t rows,
                                  const uint64_t *challenges, const uint64_t acc_in[4],
                                  uint64_t *acc_out, uint64_t *chain_out, uint64_t acc_final[4],
                                  int on_device);

/* acc[c][r] *= factors[c] for the `n_cols` column-major accumulator columns of `rows` rows: the fix-up of a row range whose
 * running products were accumulated from the neutral element, once the product of everything before the range is known
 * (the chained instance's hidden_fsm_input, ram_permutation/input.rs:52-62).  32 bytes of traffic per row and column pair
 * instead of a second pass over the encodings.  factors: host, [n_cols]. */
int zkc_scale_accumulators(zkc_ctx *ctx, uint64_t *acc, size_t n_cols, size_t rows, const uint64_t *factors, int on_device);

/* ---- records shared by circuits --------------------------------------------------------------- */
/* MemoryQuery witness, src/base_structures/memory_query/mod.rs:30-37 (64-byte record) */
typedef struct zkc_memory_query {
    uint32_t timestamp;
    uint32_t memory_page;
    uint32_t index;
    uint32_t rw_flag; /* 0/1 */
    uint32_t is_ptr;  /* 0/1 */
    uint32_t value[8]; /* little-endian u32 limbs of the U256 */
    uint32_t _pad[3];
} zkc_memory_query;

/* QueueState<F, 12>: head, tail, length (src/ram_permutation/input.rs:28-29) */
typedef struct zkc_queue_state12 {
    uint64_t head[ZKC_FULL_STATE];
    uint64_t tail[ZKC_FULL_STATE];
    uint32_t length;
    uint32_t _pad;
} zkc_queue_state12;

/* QueueState<F, 4> */
typedef struct zkc_queue_state4 {
    uint64_t head[ZKC_QUEUE_STATE];
    uint64_t tail[ZKC_QUEUE_STATE];
    uint32_t length;
    uint32_t _pad;
} zkc_queue_state4;

/* FullStateCircuitQueue::push of `n_queues` independent, initially empty memory queues of
 * `n_per_queue` records each (records laid out queue after queue): the way the reference builds
 * its inputs (src/ram_permutation/mod.rs:506-515; the push rule is restated in-repo at
 * src/main_vm/utils.rs:194-212).  prev_states (may be NULL): AoS [n_queues * n_per_queue][12], the
 * tail before each push = the second element of each FullStateCircuitQueueRawWitness entry.
 * final_states[n_queues]: head = 0, tail, length. */
int zkc_memory_queue_simulate(zkc_ctx *ctx, const zkc_memory_query *records, size_t n_per_queue,
                              size_t n_queues, uint64_t *prev_states, zkc_queue_state12 *final_states,
                              int on_device);

/* ---- ram_permutation (src/ram_permutation/mod.rs) ------------------------------------------- */
/* RamPermutationInputData, input.rs:27-31 */
typedef struct zkc_ram_input_data {
    zkc_queue_state12 unsorted_queue_initial_state;
    zkc_queue_state12 sorted_queue_initial_state;
    uint32_t non_deterministic_bootloader_memory_snapshot_length;
    uint32_t _pad;
} zkc_ram_input_data;

/* RamPermutationFSMInputOutput, input.rs:52-62 */
typedef struct zkc_ram_fsm {
    uint64_t lhs_accumulator[ZKC_NUM_REPETITIONS];
    uint64_t rhs_accumulator[ZKC_NUM_REPETITIONS];
    zkc_queue_state12 current_unsorted_queue_state;
    zkc_queue_state12 current_sorted_queue_state;
    uint32_t previous_sorting_key[3]; /* [timestamp, index, page] */
    uint32_t previous_full_key[2];    /* [index, page] */
    uint32_t previous_value[8];
    uint32_t previous_is_ptr;
    uint32_t num_nondeterministic_writes;
    uint32_t _pad;
} zkc_ram_fsm;

/* ClosedFormInputWitness<F, RamPermutationFSMInputOutput, RamPermutationInputData, ()>,
 * src/fsm_input_output/mod.rs:32-48 */
typedef struct zkc_ram_closed_form {
    uint32_t start_flag;
    uint32_t completion_flag; /* output (ignored on input, alloc_ignoring_outputs) */
    zkc_ram_input_data observable_input;
    zkc_ram_fsm hidden_fsm_input;
    zkc_ram_fsm hidden_fsm_output; /* output; on input: expected value if compare_expected != 0 */
} zkc_ram_closed_form;

/* trace columns of one loop iteration, in the order partial_accumulate_inner allocates them
 * (src/ram_permutation/mod.rs:246-381) */
enum zkc_ram_col {
    ZKC_RAM_UNSORTED_IS_EMPTY = 0, /* :247 */
    ZKC_RAM_SORTED_IS_EMPTY = 1,   /* :248 */
    ZKC_RAM_CAN_POP = 2,           /* :253 */
    ZKC_RAM_UNSORTED_ITEM = 3,     /* 13: ts, page, index, rw, is_ptr, value[8] (flatten order, memory_query/mod.rs:52-68) */
    ZKC_RAM_UNSORTED_ENC = 16,     /* 8: MemoryQuery::encode */
    ZKC_RAM_UNSORTED_HEAD = 24,    /* 12: queue head after the pop */
    ZKC_RAM_UNSORTED_LEN = 36,     /* queue length after the pop */
    ZKC_RAM_SORTED_ITEM = 37,      /* 13 */
    ZKC_RAM_SORTED_ENC = 50,       /* 8 */
    ZKC_RAM_SORTED_HEAD = 58,      /* 12 */
    ZKC_RAM_SORTED_LEN = 70,
    ZKC_RAM_TS_IS_ZERO = 71,              /* :261 */
    ZKC_RAM_PAGE_IS_BOOTLOADER_HEAP = 72, /* :263 */
    ZKC_RAM_IS_NONDET_WRITE = 73,         /* :270 */
    ZKC_RAM_NUM_NONDET_WRITES = 74,       /* :284, after the select */
    ZKC_RAM_CMP_DIFF = 75,                /* 3: limb differences of unpacked_long_comparison */
    ZKC_RAM_CMP_BORROW = 78,              /* 3: borrow out of each limb */
    ZKC_RAM_CMP_LIMB_EQ = 81,             /* 3: diff.is_zero() */
    ZKC_RAM_KEYS_EQUAL = 84,              /* :304 */
    ZKC_RAM_PREV_KEY_SMALLER = 85,        /* :304 */
    ZKC_RAM_SAME_CELL = 86,               /* :318 */
    ZKC_RAM_VALUE_EQUAL = 87,             /* :319 */
    ZKC_RAM_VALUE_IS_ZERO = 88,           /* :326 */
    ZKC_RAM_IS_ZERO = 89,                 /* :329 */
    ZKC_RAM_PTR_EQUALITY = 90,            /* :330 */
    ZKC_RAM_VALUE_AND_PTR_EQUAL = 91,     /* :331 */
    ZKC_RAM_READ_UNINIT = 92,             /* :335 / :349 (the flag is_zero is enforced under) */
    ZKC_RAM_CHECK_EQUALITY = 93,          /* :339 / :354 */
    ZKC_RAM_GP_CHAIN = 94,  /* 32: fma partials, column (rep*2 + side)*8 + i, side 0 = lhs(unsorted); utils.rs:112-128 */
    ZKC_RAM_GP_NEW = 126,   /* 4: lhs.mul / rhs.mul results, rep*2 + side; utils.rs:131-132 */
    ZKC_RAM_GP_ACC = 130,   /* 4: accumulators after the select, rep*2 + side; utils.rs:134-135 */
    /* ---- cells the gadgets called by the loop body allocate (boojum is un-vendored: what each gadget allocates is from its
     * published construction -- Num::is_zero = ZeroCheckGate (flag + inverse witness, 0 for 0), Num::equals = is_zero of
     * the difference, UInt32 / UInt256::equals = per-limb Num::equals + multi_and, decompose_into_bytes = 4 little-endian
     * bytes).  include/zkc_b200_ram_variables.json maps every column to the reference line that allocates it and lists what
     * stays host-resolved (range-check decompositions of allocated limbs, Poseidon2 round cells, linear-combination gates). */
    ZKC_RAM_UNSORTED_ENC_BYTES = 134, /* 12: decompose_into_bytes of value limbs 5, 6, 7 (memory_query/mod.rs:133-135), 4 LE bytes each */
    ZKC_RAM_SORTED_ENC_BYTES = 146,   /* 12 */
    ZKC_RAM_UNSORTED_LEN_INV = 158,   /* :247 is_empty: inverse witness of the queue length before the pop */
    ZKC_RAM_SORTED_LEN_INV = 159,     /* :248 */
    ZKC_RAM_TS_INV = 160,             /* :261 */
    ZKC_RAM_PAGE_DIFF = 161,          /* :263 UInt32::equals: memory_page - bootloader_heap_page (field) */
    ZKC_RAM_PAGE_DIFF_INV = 162,
    ZKC_RAM_CMP_DIFF_INV = 163,       /* 3: diff.is_zero() of unpacked_long_comparison, storage_validity.../mod.rs:937 */
    ZKC_RAM_CELL_DIFF = 166,          /* 2: :318 long_equals: comparison_key[i] - previous_comparison_key[i] (field) */
    ZKC_RAM_CELL_DIFF_INV = 168,      /* 2 */
    ZKC_RAM_CELL_LIMB_EQ = 170,       /* 2 */
    ZKC_RAM_VALUE_DIFF = 172,         /* 8: :319 UInt256::equals(value, previous_element_value), per limb */
    ZKC_RAM_VALUE_DIFF_INV = 180,     /* 8 */
    ZKC_RAM_VALUE_LIMB_EQ = 188,      /* 8 */
    ZKC_RAM_VALUE_ZERO_DIFF = 196,    /* 8: :326 UInt256::equals(value, zero): value[i] - 0 */
    ZKC_RAM_VALUE_ZERO_DIFF_INV = 204,/* 8 */
    ZKC_RAM_VALUE_ZERO_LIMB_EQ = 212, /* 8 */
    ZKC_RAM_PTR_DIFF = 220,           /* :330 Num::equals(previous_is_ptr, is_ptr): previous - current (field) */
    ZKC_RAM_PTR_DIFF_INV = 221,
    ZKC_RAM_NUM_COLS = 222
};

/* failed_checks bits for ram_permutation */
#define ZKC_RAM_CHK_LENGTHS_EQUAL (1u << 0)      /* :237 */
#define ZKC_RAM_CHK_EMPTY_SYNC (1u << 1)         /* :252 */
#define ZKC_RAM_CHK_ASCENDING (1u << 2)          /* :312 / :315 */
#define ZKC_RAM_CHK_UNINIT_READ_ZERO (1u << 3)   /* :336 / :351 */
#define ZKC_RAM_CHK_READ_CONSISTENT (1u << 4)    /* :340 / :356 */
#define ZKC_RAM_CHK_QUEUE_CONSISTENCY (1u << 5)  /* :161-162 */
#define ZKC_RAM_CHK_GRAND_PRODUCT (1u << 6)      /* :166-168 */
#define ZKC_RAM_CHK_NONDET_COUNT (1u << 7)       /* :170-175 */
#define ZKC_RAM_CHK_TRIVIAL_HEAD (1u << 8)       /* :58-60, :85-87 */
#define ZKC_RAM_CHK_RANGE (1u << 9)              /* allocation range checks (booleans) */
#define ZKC_RAM_CHK_QUEUE_HINT (1u << 10)        /* *_prev_states is not the hash chain of the records */

typedef struct zkc_ram_options {
    uint32_t bootloader_heap_page; /* 0 = ZKC_BOOTLOADER_HEAP_PAGE_DEFAULT */
    uint32_t compare_expected;     /* !=0: hook_compare_witness against io->hidden_fsm_output */
    uint32_t _pad[2];
} zkc_ram_options;

/* ram_permutation_entry_point, src/ram_permutation/mod.rs:31-210.
 *   unsorted/sorted         : queue witnesses in pop order (FullStateCircuitQueueRawWitness elements,
 *                             input.rs:105-116), n_* records
 *   *_prev_states           : AoS [n][12], the previous-tail element of each raw witness entry; NULL =
 *                             not supplied (the head chain is then recomputed sequentially on device)
 *   trace                   : column-major [ZKC_RAM_NUM_COLS][limit] or NULL
 *   io                      : in: start_flag, observable_input, hidden_fsm_input; out: completion_flag,
 *                             hidden_fsm_output
 *   commitment              : the 4 public inputs */
int zkc_ram_permutation_entry_point(zkc_ctx *ctx, zkc_ram_closed_form *io,
                                    const zkc_memory_query *unsorted, const uint64_t *unsorted_prev_states,
                                    size_t n_unsorted, const zkc_memory_query *sorted,
                                    const uint64_t *sorted_prev_states, size_t n_sorted, size_t limit,
                                    const zkc_ram_options *options, int on_device, uint64_t *trace,
                                    uint64_t commitment[ZKC_COMMITMENT_LEN], zkc_status *status);

/* constraint evaluation of a finished ram_permutation trace: re-evaluates every relation of the
 * loop body on every row (what the reference's test asserts through `check_if_satisfied`,
 * src/ram_permutation/mod.rs:556) and returns the number of violating rows; status->first_bad_row /
 * failed_checks (ZKC_RAMV_* bits) describe the first one.  `gates` selects the relation families:
 * ZKC_GATES_GENERAL = the streaming (HBM-bound) relations, ZKC_GATES_ROUND_FUNCTION adds the
 * Poseidon2 link of the two queue heads (integer-ALU bound); 0 = all. */
#define ZKC_GATES_GENERAL 1u
#define ZKC_GATES_ROUND_FUNCTION 2u
#define ZKC_RAMV_BOOLEAN (1u << 0)
#define ZKC_RAMV_QUEUE_LEN (1u << 1)
#define ZKC_RAMV_ENCODING (1u << 2)
#define ZKC_RAMV_ROUND_FUNCTION (1u << 3)
#define ZKC_RAMV_NONDET (1u << 4)
#define ZKC_RAMV_COMPARISON (1u << 5)
#define ZKC_RAMV_FLAGS (1u << 6)
#define ZKC_RAMV_ENFORCE (1u << 7)
#define ZKC_RAMV_GP_CHAIN (1u << 8)
#define ZKC_RAMV_GADGET_CELLS (1u << 10) /* bytes / differences / inverse witnesses of the gadget cells */
#define ZKC_RAMV_GP_ACC (1u << 9)
int zkc_ram_permutation_check_trace(zkc_ctx *ctx, const zkc_ram_closed_form *io, const uint64_t *trace,
                                    size_t limit, const zkc_ram_options *options, uint32_t gates, int on_device,
                                    uint64_t *violations, zkc_status *status);


/* ---- LogQuery circuits: log_sorter, storage_validity_by_grand_product ---------------------------- */
/* LogQuery witness, src/base_structures/log_query/mod.rs:23-35 (128-byte record).
 * flags = aux_byte | shard_id << 8 | rw_flag << 16 | rollback << 17 | is_service << 18 */
typedef struct zkc_log_query {
    uint32_t address[5];       /* UInt160, little-endian u32 limbs */
    uint32_t key[8];           /* UInt256 */
    uint32_t read_value[8];
    uint32_t written_value[8];
    uint32_t tx_number_in_block;
    uint32_t timestamp;
    uint32_t flags;
} zkc_log_query;
#define ZKC_LQ_AUX(f) ((f) & 0xFFu)
#define ZKC_LQ_SHARD(f) (((f) >> 8) & 0xFFu)
#define ZKC_LQ_RW(f) (((f) >> 16) & 1u)
#define ZKC_LQ_ROLLBACK(f) (((f) >> 17) & 1u)
#define ZKC_LQ_SERVICE(f) (((f) >> 18) & 1u)
#define ZKC_LQ_FLAGS(aux, shard, rw, rollback, service) \
    ((uint32_t)(aux) | (uint32_t)(shard) << 8 | (uint32_t)(rw) << 16 | (uint32_t)(rollback) << 17 | (uint32_t)(service) << 18)
#define ZKC_LOG_QUERY_FLAT 36   /* FLATTENED_VARIABLE_LENGTH, log_query/mod.rs:42 */
#define ZKC_LOG_QUERY_PACKED 20 /* LOG_QUERY_PACKED_WIDTH, log_query/mod.rs:38 */

/* CircuitQueue<_, LogQuery, 8, 12, 4, 4, 20, R>::push of `n_queues` independent empty queues of
 * `n_per_queue` records (StorageLogQueue, src/demux_log_queue/mod.rs:34; push rule restated in-repo at
 * src/main_vm/opcodes/log.rs:469-600: empty sponge, absorb enc[0..8], enc[8..16], enc[16..20] || tail).
 * extra_timestamps (may be NULL): per record, the TimestampedStorageLogRecord.timestamp folded into
 * element 19 (storage_validity_by_grand_product/mod.rs:72-96).  prev_tails (may be NULL): AoS [n][4], the
 * tail before each push = second element of each CircuitQueueRawWitness entry. */
int zkc_log_queue_simulate(zkc_ctx *ctx, const zkc_log_query *records, const uint32_t *extra_timestamps,
                           size_t n_per_queue, size_t n_queues, uint64_t *prev_tails,
                           zkc_queue_state4 *final_states, int on_device);

/* EventsDeduplicatorFSMInputOutput, src/log_sorter/input.rs:28-36 */
typedef struct zkc_events_fsm {
    uint64_t lhs_accumulator[ZKC_NUM_REPETITIONS];
    uint64_t rhs_accumulator[ZKC_NUM_REPETITIONS];
    zkc_queue_state4 initial_unsorted_queue_state;
    zkc_queue_state4 intermediate_sorted_queue_state;
    zkc_queue_state4 final_result_queue_state;
    uint32_t previous_key;
    uint32_t _pad;
    zkc_log_query previous_item;
} zkc_events_fsm;

/* ClosedFormInputWitness<F, EventsDeduplicatorFSMInputOutput, EventsDeduplicatorInputData,
 * EventsDeduplicatorOutputData>, src/log_sorter/input.rs:56-96 */
typedef struct zkc_events_closed_form {
    uint32_t start_flag;
    uint32_t completion_flag;                          /* out */
    zkc_queue_state4 initial_log_queue_state;          /* observable input */
    zkc_queue_state4 intermediate_sorted_queue_state;  /* observable input */
    zkc_queue_state4 final_queue_state;                /* observable output (out; expected if compared) */
    zkc_events_fsm hidden_fsm_input;
    zkc_events_fsm hidden_fsm_output;                  /* out; on input: expected value if compare_expected */
} zkc_events_closed_form;

/* trace columns of one iteration of repack_and_prove_events_rollbacks_inner (src/log_sorter/mod.rs:283-403) */
enum zkc_events_col {
    ZKC_EV_ORIGINAL_IS_EMPTY = 0, /* :284 */
    ZKC_EV_SORTED_IS_EMPTY = 1,   /* :285 */
    ZKC_EV_SHOULD_POP = 2,        /* :288 */
    ZKC_EV_UNSORTED_ITEM = 3,     /* 36, flatten order log_query/mod.rs:62-101 */
    ZKC_EV_UNSORTED_ENC = 39,     /* 20 */
    ZKC_EV_UNSORTED_HEAD = 59,    /* 4: head after the pop */
    ZKC_EV_UNSORTED_LEN = 63,
    ZKC_EV_SORTED_ITEM = 64,      /* 36 */
    ZKC_EV_SORTED_ENC = 100,      /* 20 */
    ZKC_EV_SORTED_HEAD = 120,     /* 4 */
    ZKC_EV_SORTED_LEN = 124,
    ZKC_EV_GP_CHAIN = 125,        /* 80: (rep*2 + side)*20 + i */
    ZKC_EV_GP_NEW = 205,          /* 4 */
    ZKC_EV_GP_ACC = 209,          /* 4 */
    ZKC_EV_CMP_DIFF = 213,        /* :327 unpacked_long_comparison([previous_key], [sorting_key]) */
    ZKC_EV_CMP_BORROW = 214,      /* = new_key_is_smaller */
    ZKC_EV_KEYS_EQUAL = 215,      /* same_log */
    ZKC_EV_SAME_NONTRIVIAL_LOG = 216,      /* :335 */
    ZKC_EV_DIFFERENT_NONTRIVIAL_LOG = 217, /* :337 */
    ZKC_EV_ITEM_KEYS_EQUAL = 218,          /* :353 */
    ZKC_EV_VALUES_EQUAL = 219,             /* :354 */
    ZKC_EV_SAME_BODY = 220,                /* :356 */
    ZKC_EV_PREVIOUS_IS_TRIVIAL = 221,      /* value on entry to the iteration */
    ZKC_EV_SHOULD_ENFORCE = 222,           /* :360 */
    ZKC_EV_MAYBE_ADD = 223,                /* :370 */
    ZKC_EV_ADD_TO_QUEUE = 224,             /* :372 */
    ZKC_EV_PUSH_ENC = 225,                 /* 20: encoding of query_to_add */
    ZKC_EV_PUSH_ROUND0 = 245,              /* 12: sponge state after absorbing enc[0..8] */
    ZKC_EV_PUSH_ROUND1 = 257,              /* 12 */
    ZKC_EV_PUSH_ROUND2 = 269,              /* 12: after absorbing enc[16..20] || old tail */
    ZKC_EV_RESULT_TAIL = 281,              /* 4: result queue tail after the conditional push */
    ZKC_EV_RESULT_LEN = 285,
    ZKC_EV_NUM_COLS = 286
};

#define ZKC_EV_CHK_LENGTHS_EQUAL (1u << 0)     /* :273-277 */
#define ZKC_EV_CHK_EMPTY_SYNC (1u << 1)        /* :286 and entry point :186 */
#define ZKC_EV_CHK_UNSORTED_IS_WRITE (1u << 2) /* :295-297 */
#define ZKC_EV_CHK_SORTED_IS_WRITE (1u << 3)   /* :318-320 */
#define ZKC_EV_CHK_ORDER (1u << 4)             /* :331 */
#define ZKC_EV_CHK_NOT_ROLLBACK (1u << 5)      /* :342-343 */
#define ZKC_EV_CHK_IS_ROLLBACK (1u << 6)       /* :347-349 */
#define ZKC_EV_CHK_SAME_BODY (1u << 7)         /* :362 */
#define ZKC_EV_CHK_QUEUE_CONSISTENCY (1u << 8) /* :437-438 */
#define ZKC_EV_CHK_GRAND_PRODUCT (1u << 9)     /* :189-191 */
#define ZKC_EV_CHK_TRIVIAL_HEAD (1u << 10)     /* :61, :87 */
#define ZKC_EV_CHK_QUEUE_HINT (1u << 11)       /* *_prev_tails / result_tails is not the hash chain */

typedef struct zkc_sorter_options {
    uint32_t compare_expected; /* hook_compare_witness against the expected outputs in *io */
    uint32_t _pad[3];
} zkc_sorter_options;

/* sort_and_deduplicate_events_entry_point, src/log_sorter/mod.rs:34-232.
 *   unsorted / sorted, *_prev_tails : queue witnesses in pop order (CircuitQueueRawWitness, input.rs:98-106)
 *   result_tails : NULL, or AoS [pushes][4]: the result-queue tail after each executed push (verified);
 *                  when NULL the chain is rebuilt sequentially on the device (1 permutation per push)
 *   trace        : column-major [ZKC_EV_NUM_COLS][limit] or NULL */
int zkc_log_sorter_entry_point(zkc_ctx *ctx, zkc_events_closed_form *io, const zkc_log_query *unsorted,
                               const uint64_t *unsorted_prev_tails, size_t n_unsorted, const zkc_log_query *sorted,
                               const uint64_t *sorted_prev_tails, size_t n_sorted, const uint64_t *result_tails,
                               size_t n_result_tails, size_t limit, const zkc_sorter_options *options, int on_device,
                               uint64_t *trace, uint64_t commitment[ZKC_COMMITMENT_LEN], zkc_status *status);

/* constraint evaluation of a finished log_sorter trace (as zkc_ram_permutation_check_trace): re-evaluates every relation of
 * the loop body on every row and returns the number of violating rows; status->first_bad_row / failed_checks (ZKC_EVV_* bits)
 * describe the first one.  io: start_flag, the observable input and hidden_fsm_input of the instance the trace belongs to.
 * gates: ZKC_GATES_GENERAL = the streaming (HBM-bound) relations, ZKC_GATES_ROUND_FUNCTION adds the three permutations of the
 * result-queue push; 0 = all. */
#define ZKC_EVV_BOOLEAN (1u << 0)      /* booleans, u32 / u8 ranges, field range of hash outputs */
#define ZKC_EVV_QUEUE_LEN (1u << 1)    /* is_empty / length / head bookkeeping of the two popped queues */
#define ZKC_EVV_ENCODING (1u << 2)     /* LogQuery::encode of the popped items and of the pushed record */
#define ZKC_EVV_ROUND_FUNCTION (1u << 3)
#define ZKC_EVV_COMPARISON (1u << 4)   /* :327 borrow chain */
#define ZKC_EVV_FLAGS (1u << 5)        /* :335-372 */
#define ZKC_EVV_ENFORCE (1u << 6)      /* conditional enforcements */
#define ZKC_EVV_GP_CHAIN (1u << 7)
#define ZKC_EVV_GP_ACC (1u << 8)
#define ZKC_EVV_RESULT_QUEUE (1u << 9) /* result queue length / tail selection */
int zkc_log_sorter_check_trace(zkc_ctx *ctx, const zkc_events_closed_form *io, const uint64_t *trace, size_t limit, uint32_t gates,
                               int on_device, uint64_t *violations, zkc_status *status);


/* ---- storage_validity_by_grand_product (src/storage_validity_by_grand_product/mod.rs) ------------ */
#define ZKC_PACKED_KEY_LENGTH 13 /* PACKED_KEY_LENGTH, input.rs:28 */

/* StorageDeduplicatorFSMInputOutput, input.rs:37-52 */
typedef struct zkc_storage_fsm {
    uint64_t lhs_accumulator[ZKC_NUM_REPETITIONS];
    uint64_t rhs_accumulator[ZKC_NUM_REPETITIONS];
    zkc_queue_state4 current_unsorted_queue_state;
    zkc_queue_state4 current_intermediate_sorted_queue_state;
    zkc_queue_state4 current_final_sorted_queue_state;
    uint32_t cycle_idx;
    uint32_t previous_packed_key[ZKC_PACKED_KEY_LENGTH];
    uint32_t previous_key[8];
    uint32_t previous_address[5];
    uint32_t previous_timestamp;
    uint32_t this_cell_has_explicit_read_and_rollback_depth_zero;
    uint32_t this_cell_base_value[8];
    uint32_t this_cell_current_value[8];
    uint32_t this_cell_current_depth;
    uint32_t _pad;
} zkc_storage_fsm;

/* ClosedFormInputWitness<F, StorageDeduplicatorFSMInputOutput, StorageDeduplicatorInputData,
 * StorageDeduplicatorOutputData>, input.rs:90-126 */
typedef struct zkc_storage_closed_form {
    uint32_t start_flag;
    uint32_t completion_flag;                          /* out */
    uint32_t shard_id_to_process;                      /* observable input */
    uint32_t _pad;
    zkc_queue_state4 unsorted_log_queue_state;         /* observable input */
    zkc_queue_state4 intermediate_sorted_queue_state;  /* observable input */
    zkc_queue_state4 final_sorted_queue_state;         /* observable output */
    zkc_storage_fsm hidden_fsm_input;
    zkc_storage_fsm hidden_fsm_output;
} zkc_storage_closed_form;

/* trace columns of one iteration of sort_and_deduplicate_storage_access_inner (mod.rs:584-833) */
enum zkc_storage_col {
    ZKC_ST_ORIGINAL_IS_EMPTY = 0,
    ZKC_ST_SORTED_IS_EMPTY = 1,
    ZKC_ST_SHOULD_POP = 2,
    ZKC_ST_ORIGINAL_TIMESTAMP = 3, /* cycle_idx before the increment, :585 */
    ZKC_ST_UNSORTED_ITEM = 4,      /* 36 */
    ZKC_ST_UNSORTED_ENC = 40,      /* 20: LogQuery::encode */
    ZKC_ST_UNSORTED_EXT19 = 60,    /* element 19 after append_timestamp_to_raw_query_encoding, :605-610 */
    ZKC_ST_UNSORTED_HEAD = 61,     /* 4 */
    ZKC_ST_UNSORTED_LEN = 65,
    ZKC_ST_SORTED_ITEM = 66,       /* 37: record flatten + timestamp */
    ZKC_ST_SORTED_ENC = 103,       /* 20: timestamped encoding */
    ZKC_ST_SORTED_HEAD = 123,      /* 4 */
    ZKC_ST_SORTED_LEN = 127,
    ZKC_ST_SHARD_ID_IS_VALID = 128,
    ZKC_ST_GP_CHAIN = 129,         /* 80 */
    ZKC_ST_GP_NEW = 209,           /* 4 */
    ZKC_ST_GP_ACC = 213,           /* 4 */
    ZKC_ST_CMP_DIFF = 217,         /* 13: unpacked_long_comparison(previous_packed_key, packed_key), :635-636 */
    ZKC_ST_CMP_BORROW = 230,       /* 13 */
    ZKC_ST_CMP_LIMB_EQ = 243,      /* 13 */
    ZKC_ST_KEYS_ARE_EQUAL = 256,
    ZKC_ST_PREVIOUS_KEY_IS_GREATER = 257,
    ZKC_ST_TS_DIFF = 258,          /* previous_timestamp - timestamp, :645 */
    ZKC_ST_PREVIOUS_TIMESTAMP_IS_LESS = 259,
    ZKC_ST_MUST_ENFORCE = 260,
    ZKC_ST_VALUE_IS_UNCHANGED = 261,
    ZKC_ST_CURRENT_DEPTH_IS_ZERO = 262,
    ZKC_ST_UNCHANGED_BUT_NOT_BY_ROLLBACK = 263,
    ZKC_ST_ISSUE_PROTECTIVE_READ = 264,
    ZKC_ST_SHOULD_WRITE = 265,
    ZKC_ST_SHOULD_UPDATE = 266,
    ZKC_ST_SHOULD_PUSH = 267,
    ZKC_ST_NEW_NON_TRIVIAL_CELL = 268,
    ZKC_ST_PUSH_ENC = 269,         /* 20 */
    ZKC_ST_PUSH_ROUND0 = 289,      /* 12 */
    ZKC_ST_PUSH_ROUND1 = 301,      /* 12 */
    ZKC_ST_PUSH_ROUND2 = 313,      /* 12 */
    ZKC_ST_RESULT_TAIL = 325,      /* 4 */
    ZKC_ST_RESULT_LEN = 329,
    ZKC_ST_CELL_BASE_VALUE = 330,      /* 8: this_cell_base_value at the end of the iteration */
    ZKC_ST_CELL_CURRENT_VALUE = 338,   /* 8 */
    ZKC_ST_CELL_CURRENT_DEPTH = 346,
    ZKC_ST_CELL_HAS_READ_AT_DEPTH_ZERO = 347,
    ZKC_ST_NON_TRIVIAL_AND_SAME_CELL = 348,
    ZKC_ST_READ_OF_SAME_CELL = 349,
    ZKC_ST_WRITE_OF_SAME_CELL = 350,
    ZKC_ST_WRITE_NO_ROLLBACK = 351,
    ZKC_ST_WRITE_ROLLBACK = 352,
    ZKC_ST_READ_IS_EQUAL_TO_CURRENT = 353,
    ZKC_ST_CHECK_READ_CONSISTENCY = 354,
    ZKC_ST_ROLLBACK_DEPTH_IS_ZERO = 355,
    ZKC_ST_READ_AT_DEPTH_ZERO_OF_SAME_CELL = 356,
    ZKC_ST_NUM_COLS = 357
};

#define ZKC_ST_CHK_LENGTHS_EQUAL (1u << 0)     /* :565-569 */
#define ZKC_ST_CHK_EMPTY_SYNC (1u << 1)        /* :594 and entry point :431 */
#define ZKC_ST_CHK_SHARD_ID (1u << 2)          /* :612-614 */
#define ZKC_ST_CHK_KEY_ORDER (1u << 3)         /* :638-639 */
#define ZKC_ST_CHK_TIMESTAMP_ORDER (1u << 4)   /* :645-648 */
#define ZKC_ST_CHK_FIRST_KEY_NONZERO (1u << 5) /* :657-661 */
#define ZKC_ST_CHK_READ_CONSISTENCY (1u << 6)  /* :787-792 */
#define ZKC_ST_CHK_QUEUE_CONSISTENCY (1u << 7) /* :425-426 */
#define ZKC_ST_CHK_GRAND_PRODUCT (1u << 8)     /* :434-437 */
#define ZKC_ST_CHK_TRIVIAL_HEAD (1u << 9)      /* :213, :281 */
#define ZKC_ST_CHK_QUEUE_HINT (1u << 10)
#define ZKC_ST_CHK_DEPTH_UNDERFLOW (1u << 11)  /* a rollback at depth 0: decrement_unchecked (:774) would leave the u32
                                                  range in the reference; such inputs are not legal traces */

/* sort_and_deduplicate_storage_access_entry_point, mod.rs:166-507.  Arguments as zkc_log_sorter_entry_point;
 * sorted_timestamps[n_sorted] = TimestampedStorageLogRecord.timestamp of each sorted record (mod.rs:63-66). */
int zkc_storage_validity_entry_point(zkc_ctx *ctx, zkc_storage_closed_form *io, const zkc_log_query *unsorted,
                                     const uint64_t *unsorted_prev_tails, size_t n_unsorted,
                                     const zkc_log_query *sorted, const uint32_t *sorted_timestamps,
                                     const uint64_t *sorted_prev_tails, size_t n_sorted, const uint64_t *result_tails,
                                     size_t n_result_tails, size_t limit, const zkc_sorter_options *options,
                                     int on_device, uint64_t *trace, uint64_t commitment[ZKC_COMMITMENT_LEN],
                                     zkc_status *status);


/* constraint evaluation of a finished storage_validity trace (as zkc_log_sorter_check_trace): every relation of the loop body
 * (mod.rs:560-800) on every row -- item ranges, the three encodings, queue bookkeeping, the 4 x 20 fma chains, the 13-limb key
 * and the timestamp comparison, the flag algebra, the per-cell state machine from the previous row's state, the push decision,
 * the result queue -- returns the number of violating rows; status->first_bad_row / failed_checks (ZKC_STV_* bits) describe the
 * first one.  gates: ZKC_GATES_GENERAL = the streaming (HBM-bound) relations; ZKC_GATES_ROUND_FUNCTION adds the permutations
 * (3 per popped item and queue, 3 of the push); 0 = all. */
#define ZKC_STV_BOOLEAN (1u << 0)      /* booleans, u32 / u8 ranges, field range of hash outputs, the cycle counter */
#define ZKC_STV_QUEUE_LEN (1u << 1)    /* is_empty / length / head bookkeeping of the two popped queues */
#define ZKC_STV_ENCODING (1u << 2)     /* LogQuery::encode of the popped items (+ timestamps) and of the pushed net query */
#define ZKC_STV_ROUND_FUNCTION (1u << 3)
#define ZKC_STV_COMPARISON (1u << 4)   /* :635-648 borrow chains */
#define ZKC_STV_FLAGS (1u << 5)        /* flag algebra, the push decision */
#define ZKC_STV_ENFORCE (1u << 6)      /* conditional enforcements */
#define ZKC_STV_GP_CHAIN (1u << 7)
#define ZKC_STV_GP_ACC (1u << 8)
#define ZKC_STV_RESULT_QUEUE (1u << 9) /* result queue length / tail selection */
#define ZKC_STV_CELL_STATE (1u << 10)  /* base / current value, rollback depth, explicit-read flag: the update from the previous row's state */
int zkc_storage_validity_check_trace(zkc_ctx *ctx, const zkc_storage_closed_form *io, const uint64_t *trace, size_t limit, uint32_t gates,
                                     int on_device, uint64_t *violations, zkc_status *status);


/* ---- sort_decommittment_requests (src/sort_decommittment_requests/mod.rs) ------------------------ */
/* DecommitQuery witness, src/base_structures/decommit_query/mod.rs:20-27 (48-byte record) */
typedef struct zkc_decommit_query {
    uint32_t code_hash[8]; /* UInt256, little-endian u32 limbs */
    uint32_t page;
    uint32_t is_first;     /* Boolean */
    uint32_t timestamp;
    uint32_t _pad;
} zkc_decommit_query;
#define ZKC_DECOMMIT_QUERY_FLAT 11   /* INTERNAL_STRUCT_LEN, decommit_query/mod.rs:116 */
#define ZKC_DECOMMIT_QUERY_PACKED 8  /* DECOMMIT_QUERY_PACKED_WIDTH, decommit_query/mod.rs:29 */
#define ZKC_DQ_PACKED_KEY_LENGTH 9   /* PACKED_KEY_LENGTH, sort_decommittment_requests/input.rs:21 */

/* FullStateCircuitQueue::push of `n_queues` independent, initially empty DecommitQueues (DecommitQueue,
 * decommit_query/mod.rs:174-182; the way the reference test builds its inputs, mod.rs:506-517).
 * prev_states (may be NULL): AoS [n][12], the tail before each push. */
int zkc_decommit_queue_simulate(zkc_ctx *ctx, const zkc_decommit_query *records, size_t n_per_queue, size_t n_queues,
                                uint64_t *prev_states, zkc_queue_state12 *final_states, int on_device);

/* CodeDecommittmentsDeduplicatorFSMInputOutput, input.rs:26-38 */
typedef struct zkc_decommit_sorter_fsm {
    zkc_queue_state12 initial_queue_state;
    zkc_queue_state12 sorted_queue_state;
    zkc_queue_state12 final_queue_state;
    uint64_t lhs_accumulator[ZKC_NUM_REPETITIONS];
    uint64_t rhs_accumulator[ZKC_NUM_REPETITIONS];
    uint32_t previous_packed_key[ZKC_DQ_PACKED_KEY_LENGTH]; /* [timestamp, code_hash limbs 0..7] */
    uint32_t first_encountered_timestamp;
    zkc_decommit_query previous_record;
} zkc_decommit_sorter_fsm;

/* ClosedFormInputWitness<F, CodeDecommittmentsDeduplicatorFSMInputOutput, ...InputData, ...OutputData>,
 * input.rs:63-112 */
typedef struct zkc_decommit_sorter_closed_form {
    uint32_t start_flag;
    uint32_t completion_flag;                     /* out */
    zkc_queue_state12 initial_queue_state;        /* observable input */
    zkc_queue_state12 sorted_queue_initial_state; /* observable input */
    zkc_queue_state12 final_queue_state;          /* observable output (out; expected if compared) */
    zkc_decommit_sorter_fsm hidden_fsm_input;
    zkc_decommit_sorter_fsm hidden_fsm_output;    /* out; on input: expected value if compare_expected */
} zkc_decommit_sorter_closed_form;

/* trace columns of one iteration of sort_and_deduplicate_code_decommittments_inner (mod.rs:277-355) */
enum zkc_decommit_sorter_col {
    ZKC_DQ_ORIGINAL_IS_EMPTY = 0, /* :278 */
    ZKC_DQ_SORTED_IS_EMPTY = 1,   /* :279 */
    ZKC_DQ_SHOULD_POP = 2,        /* :282 */
    ZKC_DQ_UNSORTED_ITEM = 3,     /* 11, flatten order decommit_query/mod.rs:136-152 */
    ZKC_DQ_UNSORTED_ENC = 14,     /* 8 */
    ZKC_DQ_UNSORTED_HEAD = 22,    /* 12: head after the pop */
    ZKC_DQ_UNSORTED_LEN = 34,
    ZKC_DQ_SORTED_ITEM = 35,      /* 11 */
    ZKC_DQ_SORTED_ENC = 46,       /* 8 */
    ZKC_DQ_SORTED_HEAD = 54,      /* 12 */
    ZKC_DQ_SORTED_LEN = 66,
    ZKC_DQ_GP_CHAIN = 67,         /* 32: (rep*2 + side)*8 + i */
    ZKC_DQ_GP_NEW = 99,           /* 4 */
    ZKC_DQ_GP_ACC = 103,          /* 4 */
    ZKC_DQ_CMP_DIFF = 107,        /* 9: unpacked_long_comparison(packed_key, previous_packed_key), :309-310 */
    ZKC_DQ_CMP_BORROW = 116,      /* 9; the last one = new_key_is_greater */
    ZKC_DQ_CMP_LIMB_EQ = 125,     /* 9 */
    ZKC_DQ_KEYS_ARE_EQUAL = 134,
    ZKC_DQ_SAME_HASH = 135,               /* :314 */
    ZKC_DQ_ENFORCE_MUST_BE_FIRST = 136,   /* :318 */
    ZKC_DQ_PREVIOUS_IS_TRIVIAL = 137,     /* value on entry to the iteration */
    ZKC_DQ_ENFORCE_SAME_MEMORY_PAGE = 138,/* :325-326 */
    ZKC_DQ_ADD_TO_QUEUE = 139,            /* :336 */
    ZKC_DQ_PUSH_ITEM = 140,               /* 11: record_to_add, :338-340 */
    ZKC_DQ_PUSH_ENC = 151,                /* 8 */
    ZKC_DQ_RESULT_TAIL = 159,             /* 12: result queue tail after the conditional push */
    ZKC_DQ_RESULT_LEN = 171,
    ZKC_DQ_FIRST_TIMESTAMP = 172,         /* first_encountered_timestamp after the update, :345-350 */
    ZKC_DQ_NUM_COLS = 173
};

#define ZKC_DQ_CHK_LENGTHS_EQUAL (1u << 0)     /* :265-269 */
#define ZKC_DQ_CHK_EMPTY_SYNC (1u << 1)        /* :280 and :360-362 */
#define ZKC_DQ_CHK_ORDER (1u << 2)             /* :312 */
#define ZKC_DQ_CHK_MUST_BE_FIRST (1u << 3)     /* :319-321 */
#define ZKC_DQ_CHK_SAME_MEMORY_PAGE (1u << 4)  /* :328-333 */
#define ZKC_DQ_CHK_QUEUE_CONSISTENCY (1u << 5) /* :377-378 */
#define ZKC_DQ_CHK_GRAND_PRODUCT (1u << 6)     /* :183-185 */
#define ZKC_DQ_CHK_TRIVIAL_HEAD (1u << 7)      /* :78, :93 */
#define ZKC_DQ_CHK_QUEUE_HINT (1u << 8)        /* *_prev_states / result_states is not the hash chain */

/* sort_and_deduplicate_code_decommittments_entry_point, src/sort_decommittment_requests/mod.rs:40-233.
 *   unsorted / sorted, *_prev_states : queue witnesses in pop order (FullStateCircuitQueueRawWitness, input.rs:114-131):
 *                   record + the 12-element queue state before it was pushed
 *   result_states : NULL, or AoS [pushes][12]: the result-queue tail after each executed push (verified);
 *                   when NULL the chain is rebuilt sequentially on the device (1 permutation per push)
 *   trace         : column-major [ZKC_DQ_NUM_COLS][limit] or NULL */
int zkc_sort_decommittments_entry_point(zkc_ctx *ctx, zkc_decommit_sorter_closed_form *io, const zkc_decommit_query *unsorted,
                                        const uint64_t *unsorted_prev_states, size_t n_unsorted,
                                        const zkc_decommit_query *sorted, const uint64_t *sorted_prev_states,
                                        size_t n_sorted, const uint64_t *result_states, size_t n_result_states,
                                        size_t limit, const zkc_sorter_options *options, int on_device, uint64_t *trace,
                                        uint64_t commitment[ZKC_COMMITMENT_LEN], zkc_status *status);

/* constraint evaluation of a finished sort_decommittment_requests trace (as zkc_log_sorter_check_trace): every relation of the loop
 * body (mod.rs:235-381) on every row; returns the number of violating rows, status->first_bad_row / failed_checks (ZKC_DQV_* bits)
 * describe the first one.  gates: ZKC_GATES_GENERAL = the streaming relations; ZKC_GATES_ROUND_FUNCTION adds the permutations of
 * the two pops and of the push; 0 = all. */
#define ZKC_DQV_BOOLEAN (1u << 0)      /* booleans, u32 ranges, field range of hash outputs */
#define ZKC_DQV_QUEUE_LEN (1u << 1)    /* is_empty / length / head bookkeeping of the two popped queues */
#define ZKC_DQV_ENCODING (1u << 2)     /* DecommitQuery::encode of the popped items and of the pushed record */
#define ZKC_DQV_ROUND_FUNCTION (1u << 3)
#define ZKC_DQV_COMPARISON (1u << 4)   /* :309-310 borrow chain */
#define ZKC_DQV_FLAGS (1u << 5)        /* same hash / first marker / same page / add flags, the carried first-encountered timestamp */
#define ZKC_DQV_ENFORCE (1u << 6)      /* conditional enforcements */
#define ZKC_DQV_GP_CHAIN (1u << 7)
#define ZKC_DQV_GP_ACC (1u << 8)
#define ZKC_DQV_RESULT_QUEUE (1u << 9) /* the record to add, result queue length / tail selection */
int zkc_sort_decommittments_check_trace(zkc_ctx *ctx, const zkc_decommit_sorter_closed_form *io, const uint64_t *trace, size_t limit,
                                        uint32_t gates, int on_device, uint64_t *violations, zkc_status *status);


/* ---- demux_log_queue (src/demux_log_queue/mod.rs) ------------------------------------------------- */
#define ZKC_DEMUX_NUM_QUEUES 6 /* NUM_SEPARATE_QUEUES, mod.rs:221; queue order = enum LogType, mod.rs:224-232:
                                  rollup storage, events, L1 messages, keccak256, sha256, ecrecover */

/* LogDemuxerFSMInputOutput, demux_log_queue/input.rs:24-32 */
typedef struct zkc_demux_fsm {
    zkc_queue_state4 initial_log_queue_state;
    zkc_queue_state4 output_queue_states[ZKC_DEMUX_NUM_QUEUES];
} zkc_demux_fsm;

/* ClosedFormInputWitness<F, LogDemuxerFSMInputOutput, LogDemuxerInputData, LogDemuxerOutputData>, input.rs:50-122 */
typedef struct zkc_demux_closed_form {
    uint32_t start_flag;
    uint32_t completion_flag;                                    /* out */
    zkc_queue_state4 initial_log_queue_state;                    /* observable input */
    zkc_queue_state4 output_queue_states[ZKC_DEMUX_NUM_QUEUES];  /* observable output (out; expected if compared) */
    zkc_demux_fsm hidden_fsm_input;
    zkc_demux_fsm hidden_fsm_output;                             /* out; on input: expected value if compare_expected */
} zkc_demux_closed_form;

/* the constants the loop compares with come from the un-vendored zkevm_opcode_defs (system_params): the defaults are
 * its v1.4.1 values; custom_constants != 0 replaces them (all four aux bytes distinct, all three addresses distinct) */
typedef struct zkc_demux_options {
    uint32_t compare_expected;
    uint32_t custom_constants;
    uint32_t aux_bytes[4];            /* STORAGE_AUX_BYTE 0, EVENT_AUX_BYTE 1, L1_MESSAGE_AUX_BYTE 2, PRECOMPILE_AUX_BYTE 3 */
    uint32_t precompile_addresses[3]; /* formal addresses (low limb, upper limbs zero): keccak256 0x8010, sha256 0x02, ecrecover 0x01 */
    uint32_t _pad[3];
} zkc_demux_options;

/* trace columns of one iteration of demultiplex_storage_logs_inner (mod.rs:268-393) */
enum zkc_demux_col {
    ZKC_DMX_QUEUE_IS_EMPTY = 0,  /* :271 */
    ZKC_DMX_EXECUTE = 1,         /* :272 */
    ZKC_DMX_ITEM = 2,            /* 36: popped record, flatten order log_query/mod.rs:62-101 */
    ZKC_DMX_ENC = 38,            /* 20 */
    ZKC_DMX_HEAD = 58,           /* 4: head after the pop */
    ZKC_DMX_LEN = 62,
    ZKC_DMX_IS_AUX = 63,         /* 4: is_storage / is_event / is_l1_message / is_precompile aux byte, :285-290 */
    ZKC_DMX_IS_ADDRESS = 67,     /* 3: is_keccak / is_sha256 / is_ecrecover address, :292-295 */
    ZKC_DMX_IS_ROLLUP_SHARD = 70,        /* :297 */
    ZKC_DMX_EXECUTE_PORTER_STORAGE = 71, /* :300-302, enforced false */
    ZKC_DMX_BITMASK = 72,        /* 6: execute_* per output queue, :353-360 */
    ZKC_DMX_IS_BITMASK = 78,     /* check_if_bitmask_and_if_empty, :383-384 */
    ZKC_DMX_EXEC_TAIL = 79,      /* 4: tail of the state push_with_optimize selects (:419-425) */
    ZKC_DMX_EXEC_LEN = 83,
    ZKC_DMX_PUSH_ROUND0 = 84,    /* 12: sponge state after absorbing enc[0..8] (the pop's first round has the same value) */
    ZKC_DMX_PUSH_ROUND1 = 96,    /* 12 */
    ZKC_DMX_PUSH_ROUND2 = 108,   /* 12: after absorbing enc[16..20] || selected tail: exec_queue.tail = first 4 */
    ZKC_DMX_QUEUE_TAILS = 120,   /* 24: the six output queues' tails after the iteration, :437-442 */
    ZKC_DMX_QUEUE_LENS = 144,    /* 6 */
    ZKC_DMX_NUM_COLS = 150
};

#define ZKC_DMX_CHK_TRIVIAL_HEAD (1u << 0)      /* :66-69 */
#define ZKC_DMX_CHK_PORTER_STORAGE (1u << 1)    /* :304-305 */
#define ZKC_DMX_CHK_BITMASK (1u << 2)           /* :383-384 */
#define ZKC_DMX_CHK_QUEUE_CONSISTENCY (1u << 3) /* :395 */
#define ZKC_DMX_CHK_QUEUE_HINT (1u << 4)        /* prev_tails / output_tails is not the hash chain */

/* demultiplex_storage_logs_enty_point, src/demux_log_queue/mod.rs:38-217.
 *   records, prev_tails : the log queue's witness in pop order (CircuitQueueRawWitness, input.rs:124-129)
 *   output_tails : NULL, or AoS [sum n_output_tails][4] grouped by output queue: queue q's tail after each of its
 *                  executed pushes (what the out-of-circuit demultiplexer produced; together with the initial tail these
 *                  are the `prev_tails` the six downstream circuits consume).  Verified; when NULL the six chains are
 *                  rebuilt sequentially on the device (1 permutation per push, the six queues side by side)
 *   trace        : column-major [ZKC_DMX_NUM_COLS][limit] or NULL */
int zkc_demux_log_queue_entry_point(zkc_ctx *ctx, zkc_demux_closed_form *io, const zkc_log_query *records,
                                    const uint64_t *prev_tails, size_t n_records, const uint64_t *output_tails,
                                    const size_t n_output_tails[ZKC_DEMUX_NUM_QUEUES], size_t limit,
                                    const zkc_demux_options *options, int on_device, uint64_t *trace,
                                    uint64_t commitment[ZKC_COMMITMENT_LEN], zkc_status *status);

/* constraint evaluation of a finished demux_log_queue trace (as zkc_log_sorter_check_trace): every relation of the loop body
 * (mod.rs:268-393) and of push_with_optimize (:401-447) on every row; options as for the entry point (the aux bytes / formal
 * addresses).  gates: ZKC_GATES_GENERAL = the streaming relations; ZKC_GATES_ROUND_FUNCTION adds the four permutations; 0 = all. */
#define ZKC_DMXV_BOOLEAN (1u << 0)       /* booleans, u32 / u8 ranges, field range of hash outputs */
#define ZKC_DMXV_QUEUE_LEN (1u << 1)     /* is_empty / length / head bookkeeping of the popped queue */
#define ZKC_DMXV_ENCODING (1u << 2)      /* LogQuery::encode */
#define ZKC_DMXV_ROUND_FUNCTION (1u << 3)
#define ZKC_DMXV_FLAGS (1u << 4)         /* classification flags, execute bits, the bitmask flag */
#define ZKC_DMXV_ENFORCE (1u << 5)       /* no porter storage, one class per executed row */
#define ZKC_DMXV_OUTPUT_QUEUES (1u << 6) /* selected state before the push, the six tails / lengths after it */
int zkc_demux_log_queue_check_trace(zkc_ctx *ctx, const zkc_demux_closed_form *io, const zkc_demux_options *options, const uint64_t *trace,
                                    size_t limit, uint32_t gates, int on_device, uint64_t *violations, zkc_status *status);


/* ---- keccak256_round_function (src/keccak256_round_function/mod.rs) -------------------------------- */
#define ZKC_KECCAK_RATE_BYTES 136            /* boojum KECCAK_RATE_BYTES */
#define ZKC_KECCAK_BUFFER_SIZE 192           /* KECCAK_PRECOMPILE_BUFFER_SIZE, input.rs:24 */
#define ZKC_KECCAK_MEMORY_QUERIES_PER_CYCLE 6 /* input.rs:23 */
/* zkevm_opcode_defs::system_params (un-vendored; values from memory, overridable through zkc_precompile_options) */
#define ZKC_KECCAK256_PRECOMPILE_ADDRESS_DEFAULT 0x8010u
#define ZKC_SHA256_PRECOMPILE_ADDRESS_DEFAULT 0x02u
#define ZKC_PRECOMPILE_AUX_BYTE_DEFAULT 3u

/* Keccak256RoundFunctionFSMInputOutput, input.rs:29-39 + 56-61 */
typedef struct zkc_keccak_fsm {
    uint32_t read_precompile_call;
    uint32_t read_unaligned_words_for_round;
    uint32_t padding_round;
    uint32_t completed;
    uint8_t keccak_internal_state[200]; /* [i][j][byte]: (i * 5 + j) * 8 + b; lane (x = i, y = j), little-endian bytes */
    uint32_t timestamp_to_use_for_read;
    uint32_t timestamp_to_use_for_write;
    /* Keccak256PrecompileCallParams, mod.rs:45-54 */
    uint32_t input_page;
    uint32_t input_memory_byte_offset;
    uint32_t input_memory_byte_length;
    uint32_t output_page;
    uint32_t output_word_offset;
    uint32_t needs_full_padding_round;
    /* ByteBuffer<F, 192>, buffer/mod.rs:8-11 */
    uint8_t buffer_bytes[ZKC_KECCAK_BUFFER_SIZE];
    uint32_t buffer_filled;
    uint32_t _pad;
    zkc_queue_state4 log_queue_state;
    zkc_queue_state12 memory_queue_state;
} zkc_keccak_fsm;

/* ClosedFormInputWitness<F, Keccak256RoundFunctionFSMInputOutput, PrecompileFunctionInputData,
 * PrecompileFunctionOutputData>, input.rs:78-89, base_structures/precompile_input_outputs/mod.rs:23-46 */
typedef struct zkc_keccak_closed_form {
    uint32_t start_flag;
    uint32_t completion_flag;                     /* out */
    zkc_queue_state4 initial_log_queue_state;     /* observable input */
    zkc_queue_state12 initial_memory_queue_state; /* observable input */
    zkc_queue_state12 final_memory_state;         /* observable output */
    zkc_keccak_fsm hidden_fsm_input;
    zkc_keccak_fsm hidden_fsm_output;
} zkc_keccak_closed_form;

/* trace columns of one iteration of keccak256_precompile_inner's main work cycle (mod.rs:230-665) */
enum zkc_keccak_col {
    ZKC_KC_FLAGS_IN = 0,      /* 4: read_precompile_call, read_unaligned_words_for_round, padding_round, completed on entry */
    ZKC_KC_CALL_ITEM = 4,     /* 36: popped precompile call (LogQuery flatten), :260 */
    ZKC_KC_REQ_HEAD = 40,     /* 4: requests queue head after the pop */
    ZKC_KC_REQ_LEN = 44,
    ZKC_KC_PARAMS = 45,       /* 6: precompile_call_params after the select (:294-299): input_page, byte_offset, byte_length,
                                 output_page, output_word_offset, needs_full_padding_round */
    ZKC_KC_TS_READ = 51,
    ZKC_KC_TS_WRITE = 52,
    ZKC_KC_RESET_BUFFER = 53,           /* :321 */
    ZKC_KC_READ_ZERO_LENGTH = 54,       /* have_read_zero_length_call, :325 */
    ZKC_KC_READ_NON_ZERO_LENGTH = 55,   /* :331 */
    ZKC_KC_QUERY = 56,        /* 6 x 28: aligned_index, unalignment, meaningful_bytes, should_read, value[8], memory tail[12],
                                 memory length, byte_offset after, byte_length after, buffer filled after (:397-492) */
    ZKC_KC_QUERY_STRIDE = 28,
    ZKC_KC_ZERO_BYTES_LEFT = 224,       /* :497 */
    ZKC_KC_CURRENTLY_FILLED = 225,      /* :502 */
    ZKC_KC_DO_ONE_BYTE_OF_PADDING = 226,
    ZKC_KC_BUFFER_NOW_EMPTY = 227,      /* :510 */
    ZKC_KC_APPLY_PADDING = 228,         /* :515 */
    ZKC_KC_INPUT = 229,       /* 136: the absorbed block after padding / full-padding select, :580 */
    ZKC_KC_STATE_OUT = 365,   /* 200: keccak_internal_state after the permutation */
    ZKC_KC_WRITE_RESULT = 565,
    ZKC_KC_RESULT = 566,      /* 8: UInt256::from_be_bytes(squeezed) */
    ZKC_KC_WRITE_TAIL = 574,  /* 12: memory queue tail after the conditional digest write */
    ZKC_KC_WRITE_LEN = 586,
    ZKC_KC_FLAGS_OUT = 587,   /* 4 */
    ZKC_KC_BUFFER_OUT = 591,  /* 192: buffer bytes after the cycle */
    ZKC_KC_NUM_COLS = 783
};

#define ZKC_KC_CHK_TRIVIAL_HEAD (1u << 0)      /* :705, :719 */
#define ZKC_KC_CHK_AUX_BYTE (1u << 1)          /* :262-267 */
#define ZKC_KC_CHK_ADDRESS (1u << 2)           /* :268-280 */
#define ZKC_KC_CHK_BUFFER_OVERFLOW (1u << 3)   /* add_no_overflow / sub_no_overflow in the buffer, buffer/mod.rs:132-135 */
#define ZKC_KC_CHK_QUEUE_CONSISTENCY (1u << 4) /* :667 */
#define ZKC_KC_CHK_QUEUE_HINT (1u << 5)
#define ZKC_KC_CHK_WITNESS_EXHAUSTED (1u << 6) /* `expect("not empty witness")`, storage_application/mod.rs:129-131 */

typedef struct zkc_precompile_options {
    uint32_t compare_expected;
    uint32_t precompile_address; /* low 32 bits of the formal address (upper limbs 0); 0 = circuit default */
    uint32_t aux_byte;           /* 0 = ZKC_PRECOMPILE_AUX_BYTE_DEFAULT */
    uint32_t _pad;
} zkc_precompile_options;

/* constraint evaluation of a finished keccak256_round_function trace: the FSM state on entry to a cycle is rebuilt from the
 * previous cycle's cells, the cycle function (mod.rs:230-665) is re-run on the cycle's free inputs (popped call, words read) and
 * every cell it produces is compared with the trace; plus the conditional pop and the memory queue's length / tail chain over the
 * seven conditional pushes.  gates: ZKC_GATES_GENERAL = everything but the Poseidon2 permutations of the queues;
 * ZKC_GATES_ROUND_FUNCTION adds them; 0 = all. */
#define ZKC_KCV_BOOLEAN (1u << 0)        /* booleans, u32 / u8 ranges, field range of hash outputs, a call item where nothing is popped */
#define ZKC_KCV_QUEUE (1u << 1)          /* length / head bookkeeping of the requests queue */
#define ZKC_KCV_FSM (1u << 2)            /* flags on entry / exit, reset / read flags, padding decisions, write_result */
#define ZKC_KCV_ROUND_FUNCTION (1u << 3)
#define ZKC_KCV_PARAMS (1u << 4)         /* call parameters / timestamps after the selects */
#define ZKC_KCV_ENFORCE (1u << 5)        /* aux byte / formal address of a popped call; buffer fill within range */
#define ZKC_KCV_SPONGE (1u << 6)         /* absorbed block, keccak state after the permutation, digest word */
#define ZKC_KCV_MEMORY_QUEUE (1u << 7)   /* memory queue length / tail over the seven conditional pushes */
#define ZKC_KCV_QUERIES (1u << 8)        /* the six read slots: aligned index, unalignment, meaningful bytes, should_read, value, offsets, fill */
#define ZKC_KCV_BUFFER (1u << 9)         /* buffer bytes after the cycle */
int zkc_keccak256_round_function_check_trace(zkc_ctx *ctx, const zkc_keccak_closed_form *io, const zkc_precompile_options *options,
                                             const uint64_t *trace, size_t limit, uint32_t gates, int on_device, uint64_t *violations,
                                             zkc_status *status);

/* keccak256_round_function_entry_point, mod.rs:673-794.
 *   requests, requests_prev_tails : precompile calls queue witness in pop order (CircuitQueueRawWitness, input.rs:97)
 *   memory_reads                  : memory_reads_witness (input.rs:98), n_reads x 8 little-endian u32 limbs, FIFO order
 *   memory_states                 : NULL, or AoS [pushes][12]: memory queue tail after each executed push (verified);
 *                                   when NULL the chain is rebuilt sequentially on the device
 *   trace                         : column-major [ZKC_KC_NUM_COLS][limit] or NULL */
int zkc_keccak256_round_function_entry_point(zkc_ctx *ctx, zkc_keccak_closed_form *io, const zkc_log_query *requests,
                                             const uint64_t *requests_prev_tails, size_t n_requests,
                                             const uint32_t *memory_reads, size_t n_reads, const uint64_t *memory_states,
                                             size_t n_memory_states, size_t limit, const zkc_precompile_options *options,
                                             int on_device, uint64_t *trace, uint64_t commitment[ZKC_COMMITMENT_LEN],
                                             zkc_status *status);


/* ---- sha256_round_function (src/sha256_round_function/mod.rs) --------------------------------------- */
/* Sha256RoundFunctionFSMInputOutput, input.rs:23-45 */
typedef struct zkc_sha256_fsm {
    uint32_t read_precompile_call;
    uint32_t read_words_for_round;
    uint32_t completed;
    uint32_t sha256_inner_state[8];
    uint32_t timestamp_to_use_for_read;
    uint32_t timestamp_to_use_for_write;
    /* Sha256PrecompileCallParams, mod.rs:44-50 */
    uint32_t input_page;
    uint32_t input_offset;
    uint32_t output_page;
    uint32_t output_offset;
    uint32_t num_rounds;
    uint32_t _pad;
    zkc_queue_state4 log_queue_state;
    zkc_queue_state12 memory_queue_state;
} zkc_sha256_fsm;

typedef struct zkc_sha256_closed_form {
    uint32_t start_flag;
    uint32_t completion_flag;                     /* out */
    zkc_queue_state4 initial_log_queue_state;     /* observable input */
    zkc_queue_state12 initial_memory_queue_state; /* observable input */
    zkc_queue_state12 final_memory_state;         /* observable output */
    zkc_sha256_fsm hidden_fsm_input;
    zkc_sha256_fsm hidden_fsm_output;
} zkc_sha256_closed_form;

/* trace columns of one iteration of sha256_precompile_inner's main work cycle (mod.rs:146-330) */
enum zkc_sha256_col {
    ZKC_SH_FLAGS_IN = 0,     /* 3: read_precompile_call, read_words_for_round, completed on entry */
    ZKC_SH_CALL_ITEM = 3,    /* 36 */
    ZKC_SH_REQ_HEAD = 39,    /* 4 */
    ZKC_SH_REQ_LEN = 43,
    ZKC_SH_PARAMS = 44,      /* 5: after the select (:180-185): input_page, input_offset, output_page, output_offset, num_rounds */
    ZKC_SH_TS_READ = 49,
    ZKC_SH_TS_WRITE = 50,
    ZKC_SH_RESET_BUFFER = 51, /* :204 */
    ZKC_SH_SHOULD_READ = 52,  /* :215 */
    ZKC_SH_QUERY = 53,        /* 2 x 22: value[8], memory tail[12], memory length, input_offset after (:219-247) */
    ZKC_SH_QUERY_STRIDE = 22,
    ZKC_SH_MESSAGE = 97,      /* 16 big-endian message words (:250-254) */
    ZKC_SH_NUM_ROUNDS = 113,  /* after the conditional decrement, :257-268 */
    ZKC_SH_STATE_IN = 114,    /* 8: state the compression starts from (IV on reset), :271-278 */
    ZKC_SH_STATE_OUT = 122,   /* 8 */
    ZKC_SH_WRITE_RESULT = 130,
    ZKC_SH_RESULT = 131,      /* 8: write_word limbs (:290-299) */
    ZKC_SH_WRITE_TAIL = 139,  /* 12 */
    ZKC_SH_WRITE_LEN = 151,
    ZKC_SH_FLAGS_OUT = 152,   /* 3 */
    ZKC_SH_NUM_COLS = 155
};
#define ZKC_SH_CHK_ZERO_ROUNDS (1u << 7) /* a call with num_rounds = 0: decrement_unchecked (:257-262) would leave the u32
                                            range in the reference; not a legal call (other bits as ZKC_KC_CHK_*) */

/* sha256_round_function_entry_point, mod.rs:343-470.  Arguments as zkc_keccak256_round_function_entry_point. */
int zkc_sha256_round_function_entry_point(zkc_ctx *ctx, zkc_sha256_closed_form *io, const zkc_log_query *requests,
                                          const uint64_t *requests_prev_tails, size_t n_requests,
                                          const uint32_t *memory_reads, size_t n_reads, const uint64_t *memory_states,
                                          size_t n_memory_states, size_t limit, const zkc_precompile_options *options,
                                          int on_device, uint64_t *trace, uint64_t commitment[ZKC_COMMITMENT_LEN],
                                          zkc_status *status);

/* constraint evaluation of a finished sha256_round_function trace: every relation of sha256_precompile_inner (mod.rs:146-330) on
 * every cycle, the SHA-256 compression included; options as for the entry point (aux byte / formal address).  gates:
 * ZKC_GATES_GENERAL = everything but the Poseidon2 permutations of the queues; ZKC_GATES_ROUND_FUNCTION adds them; 0 = all. */
#define ZKC_SHV_BOOLEAN (1u << 0)        /* booleans, u32 / u8 ranges, field range of hash outputs, zero values when nothing is read */
#define ZKC_SHV_QUEUE (1u << 1)          /* length / head bookkeeping of the requests queue */
#define ZKC_SHV_FSM (1u << 2)            /* FSM flags carried between cycles, reset / should_read / write_result, the next flags */
#define ZKC_SHV_ROUND_FUNCTION (1u << 3)
#define ZKC_SHV_PARAMS (1u << 4)         /* call parameters / timestamps after the selects, offset increments, the round counter */
#define ZKC_SHV_ENFORCE (1u << 5)        /* aux byte / formal address of a popped call */
#define ZKC_SHV_COMPRESSION (1u << 6)    /* message words, starting state, SHA-256 compression, result word */
#define ZKC_SHV_MEMORY_QUEUE (1u << 7)   /* memory queue length / tail over the three conditional pushes */
int zkc_sha256_round_function_check_trace(zkc_ctx *ctx, const zkc_sha256_closed_form *io, const zkc_precompile_options *options,
                                          const uint64_t *trace, size_t limit, uint32_t gates, int on_device, uint64_t *violations,
                                          zkc_status *status);


/* ---- code_unpacker_sha256 (src/code_unpacker_sha256/mod.rs) --------------------------------------------- */
/* CodeDecommittmentFSM, code_unpacker_sha256/input.rs:27-38 */
typedef struct zkc_code_decommittment_fsm {
    uint32_t sha256_inner_state[8];
    uint32_t hash_to_compare_against[8]; /* UInt256, little-endian u32 limbs (limb 7 cleared) */
    uint32_t current_index;
    uint32_t current_page;
    uint32_t timestamp;
    uint32_t num_rounds_left;            /* UInt16 */
    uint32_t length_in_bits;
    uint32_t state_get_from_queue;
    uint32_t state_decommit;
    uint32_t finished;
} zkc_code_decommittment_fsm;

/* CodeDecommitterFSMInputOutput, input.rs:70-74 */
typedef struct zkc_code_unpacker_fsm {
    zkc_code_decommittment_fsm internal_fsm;
    zkc_queue_state12 decommittment_requests_queue_state;
    zkc_queue_state12 memory_queue_state;
} zkc_code_unpacker_fsm;

/* ClosedFormInputWitness<F, CodeDecommitterFSMInputOutput, CodeDecommitterInputData, CodeDecommitterOutputData>,
 * input.rs:92-150 */
typedef struct zkc_code_unpacker_closed_form {
    uint32_t start_flag;
    uint32_t completion_flag;                              /* out */
    zkc_queue_state12 memory_queue_initial_state;          /* observable input */
    zkc_queue_state12 sorted_requests_queue_initial_state; /* observable input */
    zkc_queue_state12 memory_queue_final_state;            /* observable output (out; expected if compared) */
    zkc_code_unpacker_fsm hidden_fsm_input;
    zkc_code_unpacker_fsm hidden_fsm_output;               /* out; on input: expected value if compare_expected */
} zkc_code_unpacker_closed_form;

#define ZKC_CODE_HASH_VERSION_TOP16 0x0100u /* ContractCodeSha256::VERSION_BYTE << 8 (zkevm_opcode_defs, un-vendored), mod.rs:187-189 */

/* trace columns of one iteration of unpack_code_into_memory_inner (mod.rs:191-447) */
enum zkc_code_unpacker_col {
    ZKC_CU_FLAGS_IN = 0,          /* 3: state_get_from_queue, state_decommit, finished on entry */
    ZKC_CU_REQUEST = 3,           /* 11: popped DecommitQuery (zero if nothing popped), :195-196 */
    ZKC_CU_REQ_HEAD = 14,         /* 12: requests queue head after the pop */
    ZKC_CU_REQ_LEN = 26,
    ZKC_CU_VERSION_MATCHES = 27,  /* :202 */
    ZKC_CU_LENGTH_IN_WORDS = 28,  /* :207-213 */
    ZKC_CU_LENGTH_IN_ROUNDS = 29, /* :215-221 */
    ZKC_CU_LENGTH_IN_BITS = 30,   /* state.length_in_bits after the selection, :240-245 */
    ZKC_CU_TIMESTAMP = 31,
    ZKC_CU_PAGE = 32,
    ZKC_CU_HASH_TO_COMPARE = 33,  /* 8 */
    ZKC_CU_DECOMMIT = 41,         /* state_decommit || state_get_from_queue, :277 */
    ZKC_CU_NUM_ROUNDS_LEFT = 42,  /* after the conditional decrement, :281-287 */
    ZKC_CU_LAST_ROUND = 43,
    ZKC_CU_FINALIZE = 44,
    ZKC_CU_PROCESS_SECOND_WORD = 45,
    ZKC_CU_WORD0 = 46,            /* 8: code_word_0, little-endian u32 limbs */
    ZKC_CU_WORD1 = 54,            /* 8 */
    ZKC_CU_INDEX0 = 62,           /* index of mem_query_0 */
    ZKC_CU_INDEX1 = 63,           /* index of mem_query_1 */
    ZKC_CU_INDEX_OUT = 64,        /* state.current_index after the cycle */
    ZKC_CU_MEM_TAIL0 = 65,        /* 12 + length: memory queue after the first conditional push, :351 */
    ZKC_CU_MEM_TAIL1 = 78,        /* 12 + length: after the second, :352 */
    ZKC_CU_MESSAGE = 91,          /* 16: sha256 block as big-endian words, padding selected in on finalize, :354-381 */
    ZKC_CU_STATE_IN = 107,        /* 8 */
    ZKC_CU_STATE_NEW = 115,       /* 8: round function output, :383-384 */
    ZKC_CU_STATE_OUT = 123,       /* 8: state.sha256_inner_state after the selection, :386-391 */
    ZKC_CU_FLAGS_OUT = 131,       /* 3 */
    ZKC_CU_NUM_COLS = 134
};

#define ZKC_CU_CHK_VERSION (1u << 0)           /* :202-204 */
#define ZKC_CU_CHK_LENGTH (1u << 1)            /* (length_in_words + 1) / 2 is not a UInt16, :215-221 */
#define ZKC_CU_CHK_HASH (1u << 2)              /* :409-420 */
#define ZKC_CU_CHK_QUEUE_CONSISTENCY (1u << 3) /* :449 */
#define ZKC_CU_CHK_QUEUE_HINT (1u << 4)        /* requests_prev_states / memory_states is not the hash chain */
#define ZKC_CU_CHK_WITNESS_EXHAUSTED (1u << 5) /* a pop from the empty requests queue / code_words ran dry */

/* unpack_code_into_memory_entry_point, src/code_unpacker_sha256/mod.rs:33-148.
 *   requests, requests_prev_states : sorted_requests_queue_witness (FullStateCircuitQueueRawWitness, input.rs:157-158)
 *   code_words    : `code_words` flattened (input.rs:159), n_code_words x 8 little-endian u32 limbs, in pop order
 *   memory_states : NULL, or AoS [pushes][12]: the memory queue's tail after each executed push (verified);
 *                   when NULL the chain is rebuilt sequentially on the device (1 permutation per push)
 *   trace         : column-major [ZKC_CU_NUM_COLS][limit] or NULL
 * One thread per decommitment request chains its SHA-256 rounds; requests run side by side. */
int zkc_code_unpacker_entry_point(zkc_ctx *ctx, zkc_code_unpacker_closed_form *io, const zkc_decommit_query *requests,
                                  const uint64_t *requests_prev_states, size_t n_requests, const uint32_t *code_words,
                                  size_t n_code_words, const uint64_t *memory_states, size_t n_memory_states, size_t limit,
                                  const zkc_sorter_options *options, int on_device, uint64_t *trace,
                                  uint64_t commitment[ZKC_COMMITMENT_LEN], zkc_status *status);

/* constraint evaluation of a finished code_unpacker_sha256 trace: every relation of unpack_code_into_memory_inner (mod.rs:191-447)
 * on every cycle, the SHA-256 compression and the hash comparison included.  gates: ZKC_GATES_GENERAL = everything but the
 * Poseidon2 permutations of the queues; ZKC_GATES_ROUND_FUNCTION adds them; 0 = all. */
#define ZKC_CUV_BOOLEAN (1u << 0)        /* booleans, u32 ranges, field range of hash outputs, zero words when not taken */
#define ZKC_CUV_QUEUE (1u << 1)          /* length / head bookkeeping of the requests queue */
#define ZKC_CUV_FSM (1u << 2)            /* FSM flags carried between cycles, decommit, round counter, phase flags, next flags */
#define ZKC_CUV_ROUND_FUNCTION (1u << 3)
#define ZKC_CUV_LENGTH (1u << 4)         /* version match, length in words / rounds */
#define ZKC_CUV_ENFORCE (1u << 5)        /* version of a popped request; digest = requested hash on finalize */
#define ZKC_CUV_COMPRESSION (1u << 6)    /* SHA-256 block, starting state, compression, state select */
#define ZKC_CUV_MEMORY_QUEUE (1u << 7)   /* memory queue length / tail over the two conditional writes */
#define ZKC_CUV_SELECTS (1u << 8)        /* length in bits, timestamp, page, hash, indices after the selects */
int zkc_code_unpacker_check_trace(zkc_ctx *ctx, const zkc_code_unpacker_closed_form *io, const uint64_t *trace, size_t limit, uint32_t gates,
                                  int on_device, uint64_t *violations, zkc_status *status);


/* ---- main_vm (src/main_vm/) ------------------------------------------------------------------------- */
/* The ISA tables (zkevm_opcode_defs, un-vendored) are INPUT DATA: opcode -> (price, 48-bit property bit spread +
 * 3 aux bits), exactly the 3-column table of src/tables/opcodes_decoding.rs:14-38, plus the condition table of
 * src/tables/conditional.rs:21-58 and the NOP / PANIC encodings the circuit masks with (main_vm/utils.rs:14-41,
 * decoded_opcode.rs:120-157).  The LAYOUT of the 38 meaningful property bits is main_vm/opcode_bitmask.rs:83-127:
 * [opcode type | 10 variant | 2 flags | 6 src addressing | 4 dst addressing]; the positions below are this
 * engine's restatement of zkevm_opcode_defs' enumeration order (from memory -- the host shim passes the real
 * tables, so only the enumeration order has to be confirmed against the crate). */
#define ZKC_VM_OPCODES_TABLE_WIDTH 11
#define ZKC_VM_NUM_OPCODES (1u << ZKC_VM_OPCODES_TABLE_WIDTH)
#define ZKC_VM_OPCODE_TYPE_BITS 16
#define ZKC_VM_VARIANT_BITS 10
#define ZKC_VM_FLAG_BITS 2
#define ZKC_VM_SRC_MODE_BITS 6
#define ZKC_VM_DST_MODE_BITS 4
#define ZKC_VM_DESCRIPTION_BITS 38
#define ZKC_VM_DESCRIPTION_BITS_FLATTENED 48
#define ZKC_VM_REGISTERS 15 /* REGISTERS_COUNT, base_structures/vm_state/mod.rs:30 */
enum zkc_vm_opcode_type { /* zkevm_opcode_defs::Opcode variant order */
    ZKC_OP_INVALID = 0, ZKC_OP_NOP, ZKC_OP_ADD, ZKC_OP_SUB, ZKC_OP_MUL, ZKC_OP_DIV, ZKC_OP_JUMP, ZKC_OP_CONTEXT,
    ZKC_OP_SHIFT, ZKC_OP_BINOP, ZKC_OP_PTR, ZKC_OP_NEAR_CALL, ZKC_OP_LOG, ZKC_OP_FAR_CALL, ZKC_OP_RET, ZKC_OP_UMA
};
enum zkc_vm_variant { /* materialize_subvariant_idx of the multi-variant opcodes */
    ZKC_VAR_CONTEXT_THIS = 0, ZKC_VAR_CONTEXT_CALLER, ZKC_VAR_CONTEXT_CODE_ADDRESS, ZKC_VAR_CONTEXT_META,
    ZKC_VAR_CONTEXT_ERGS_LEFT, ZKC_VAR_CONTEXT_SP, ZKC_VAR_CONTEXT_GET_U128, ZKC_VAR_CONTEXT_SET_U128,
    ZKC_VAR_CONTEXT_SET_ERGS_PER_PUBDATA, ZKC_VAR_CONTEXT_INC_TX_NUMBER,
    ZKC_VAR_SHIFT_SHL = 0, ZKC_VAR_SHIFT_SHR, ZKC_VAR_SHIFT_ROL, ZKC_VAR_SHIFT_ROR,
    ZKC_VAR_BINOP_XOR = 0, ZKC_VAR_BINOP_AND, ZKC_VAR_BINOP_OR,
    ZKC_VAR_PTR_ADD = 0, ZKC_VAR_PTR_SUB, ZKC_VAR_PTR_PACK, ZKC_VAR_PTR_SHRINK,
    ZKC_VAR_LOG_STORAGE_READ = 0, ZKC_VAR_LOG_STORAGE_WRITE, ZKC_VAR_LOG_TO_L1_MESSAGE, ZKC_VAR_LOG_EVENT, ZKC_VAR_LOG_PRECOMPILE_CALL,
    ZKC_VAR_RET_OK = 0, ZKC_VAR_RET_REVERT, ZKC_VAR_RET_PANIC,
    ZKC_VAR_UMA_HEAP_READ = 0, ZKC_VAR_UMA_HEAP_WRITE, ZKC_VAR_UMA_AUX_HEAP_READ, ZKC_VAR_UMA_AUX_HEAP_WRITE, ZKC_VAR_UMA_FAT_PTR_READ,
    ZKC_VAR_FAR_CALL_NORMAL = 0, ZKC_VAR_FAR_CALL_DELEGATE, ZKC_VAR_FAR_CALL_MIMIC
};
enum zkc_vm_src_mode { /* ImmMemHandlerFlags::variant_index */
    ZKC_MODE_REG_ONLY = 0, ZKC_MODE_STACK_PUSH_POP, ZKC_MODE_STACK_OFFSET, ZKC_MODE_STACK_ABSOLUTE, ZKC_MODE_IMM16, ZKC_MODE_CODE_PAGE
};
#define ZKC_VM_SET_FLAGS_FLAG_IDX 0       /* SET_FLAGS_FLAG_IDX */
#define ZKC_VM_SWAP_OPERANDS_FLAG_IDX 1   /* SWAP_OPERANDS_FLAG_IDX_FOR_ARITH_OPCODES */
#define ZKC_VM_SWAP_OPERANDS_PTR_FLAG_IDX 0 /* SWAP_OPERANDS_FLAG_IDX_FOR_PTR_OPCODE */
#define ZKC_VM_RET_TO_LABEL_FLAG_IDX 0    /* zkevm_opcode_defs::ret::RET_TO_LABEL_BIT_IDX */
#define ZKC_VM_UMA_INCREMENT_FLAG_IDX 0   /* UMA_INCREMENT_FLAG_IDX */
#define ZKC_VM_FIRST_MESSAGE_FLAG_IDX 0   /* FIRST_MESSAGE_FLAG_IDX (log.event / log.to_l1) */
/* call / ret ABI (zkevm_opcode_defs::definitions::abi, from memory of v1.4.1): the top 64 bits of src0 are
 * [ergs_passed u32 | forwarding mode byte | shard id byte | constructor call byte | system call byte] */
#define ZKC_VM_ABI_FORWARDING_MODE_BYTE_IDX 28 /* FAR_CALL_FORWARDING_MODE_BYTE_IDX; ret uses the same byte */
#define ZKC_VM_FORWARD_USE_HEAP 0              /* FarCallForwardPageType::UseHeap */
#define ZKC_VM_FORWARD_FAT_POINTER 1           /* ForwardFatPointer */
#define ZKC_VM_FORWARD_USE_AUX_HEAP 2          /* UseAuxHeap */
#define ZKC_VM_ABI_SHARD_ID_BYTE_IDX 29         /* FAR_CALL_SHARD_ID_BYTE_IDX */
#define ZKC_VM_ABI_CONSTRUCTOR_CALL_BYTE_IDX 30 /* FAR_CALL_CONSTRUCTOR_CALL_BYTE_IDX */
#define ZKC_VM_ABI_SYSTEM_CALL_BYTE_IDX 31      /* FAR_CALL_SYSTEM_CALL_BYTE_IDX */
#define ZKC_VM_FAR_CALL_STATIC_FLAG_IDX 0       /* FAR_CALL_STATIC_FLAG_IDX */
#define ZKC_VM_FAR_CALL_SHARD_FLAG_IDX 1        /* FAR_CALL_SHARD_FLAG_IDX */
#define ZKC_VM_AUX_KERNEL_MODE 0          /* KERNER_MODE_FLAG_IDX */
#define ZKC_VM_AUX_CAN_BE_USED_IN_STATIC 1
#define ZKC_VM_AUX_EXPLICIT_PANIC 2
/* bit position of a property inside the 38-bit description */
#define ZKC_VM_BIT_TYPE(t) (t)
#define ZKC_VM_BIT_VARIANT(v) (ZKC_VM_OPCODE_TYPE_BITS + (v))
#define ZKC_VM_BIT_FLAG(f) (ZKC_VM_OPCODE_TYPE_BITS + ZKC_VM_VARIANT_BITS + (f))
#define ZKC_VM_BIT_SRC_MODE(m) (ZKC_VM_OPCODE_TYPE_BITS + ZKC_VM_VARIANT_BITS + ZKC_VM_FLAG_BITS + (m))
#define ZKC_VM_BIT_DST_MODE(m) (ZKC_VM_OPCODE_TYPE_BITS + ZKC_VM_VARIANT_BITS + ZKC_VM_FLAG_BITS + ZKC_VM_SRC_MODE_BITS + (m))

typedef struct zkc_vm_isa {
    uint32_t opcode_price[ZKC_VM_NUM_OPCODES];   /* OPCODES_PRICES */
    uint64_t opcode_props[ZKC_VM_NUM_OPCODES];   /* OPCODES_PROPS_INTEGER_BITMASKS: 48 bits + 3 aux bits at bit 48 */
    uint8_t condition_table[8][8];               /* [condition][of | eq << 1 | gt << 2] -> 0/1 */
    uint64_t nop_opcode_encoding;                /* EncodingModeProduction::nop_encoding() */
    uint64_t panic_opcode_encoding;              /* exception_revert_encoding() */
    uint64_t nop_bitspread;                      /* NOP_BITSPREAD_U64 */
    uint64_t panic_bitspread;                    /* PANIC_BITSPREAD_U64 */
    /* zkevm_opcode_defs::system_params used by initial_bootloader_state (main_vm/loading.rs:13-226) and vm_cycle */
    uint32_t bootloader_base_page, bootloader_code_page, bootloader_calldata_page;
    uint32_t starting_timestamp, starting_base_page;
    uint32_t initial_frame_formal_eh_location, vm_initial_frame_ergs, bootloader_formal_address_low;
    uint32_t bootloader_max_memory, vm_max_stack_depth;
    /* log opcode constants (main_vm/opcodes/log.rs:128-167): aux bytes of storage / event / L1 message / precompile
     * queries and the pubdata byte counts a rollup storage write / an L1 message pay for */
    uint32_t log_aux_bytes[4];
    uint32_t initial_storage_write_pubdata_bytes, l1_message_pubdata_bytes;
    /* far call constants (call_ret_impl/far_call.rs): NEW_FRAME_MEMORY_STIPEND :343-351, NEW_MEMORY_PAGES_PER_FAR_CALL :447,
     * DEPLOYER_SYSTEM_CONTRACT_ADDRESS_LOW :1166, ERGS_PER_CODE_WORD_DECOMMITTMENT :1446, ContractCodeSha256 VERSION_BYTE /
     * YET_CONSTRUCTED_MARKER / CODE_AT_REST_MARKER :513-557, and the register conventions of
     * zkevm_opcode_defs::definitions::far_call (:1041-1066): CALL_SYSTEM_ABI_REGISTERS [lo, hi), CALL_RESERVED_RANGE [lo, hi),
     * CALL_IMPLICIT_PARAMETER_REG_IDX (0-based register indices) */
    uint32_t new_frame_memory_stipend, new_memory_pages_per_far_call, deployer_system_contract_address_low;
    uint32_t ergs_per_code_word_decommittment;
    uint32_t code_hash_version_byte, code_hash_yet_constructed_marker, code_hash_at_rest_marker;
    uint32_t call_system_abi_registers[2], call_reserved_range[2], call_implicit_parameter_reg_idx;
} zkc_vm_isa;

/* VMRegister, base_structures/register/mod.rs:21-24 */
typedef struct zkc_vm_register {
    uint32_t is_pointer;
    uint32_t value[8];
} zkc_vm_register;

/* FullExecutionContext = ExecutionContextRecord + forward log tail, vm_state/saved_context.rs:36-59,
 * vm_state/callstack.rs:45-49 (field order = declaration order = var-length encoding order) */
typedef struct zkc_vm_context {
    uint32_t this_address[5], caller[5], code_address[5];
    uint32_t code_page, base_page, heap_upper_bound, aux_heap_upper_bound;
    uint64_t reverted_queue_head[4], reverted_queue_tail[4];
    uint32_t reverted_queue_segment_len;
    uint32_t pc, sp, exception_handler_loc; /* u16 values */
    uint32_t ergs_remaining;
    uint32_t is_static_execution, is_kernel_mode;
    uint32_t this_shard_id, caller_shard_id, code_shard_id; /* u8 values */
    uint32_t context_u128_value_composite[4];
    uint32_t is_local_call;
    uint32_t log_queue_forward_part_length;
    uint64_t log_queue_forward_tail[4];
} zkc_vm_context;

/* VmLocalState, vm_state/mod.rs:92-109: the per-cycle snapshot (243 field elements when flattened) */
typedef struct zkc_vm_state {
    uint32_t previous_code_word[8];
    zkc_vm_register registers[ZKC_VM_REGISTERS];
    uint32_t flags[3]; /* overflow_or_less_than, equal, greater_than */
    uint32_t timestamp, memory_page_counter, tx_number_in_block, previous_code_page, previous_super_pc;
    uint32_t pending_exception, ergs_per_pubdata_byte;
    uint32_t context_stack_depth;
    uint32_t memory_queue_length, code_decommittment_queue_length;
    uint32_t context_composite_u128[4];
    uint32_t _pad;
    zkc_vm_context current_context;
    uint64_t stack_sponge_state[12];
    uint64_t memory_queue_state[12];
    uint64_t code_decommittment_queue_state[12];
} zkc_vm_state;
#define ZKC_VM_STATE_FLAT 243
#define ZKC_VM_STATE_WORDS 294   /* sizeof(zkc_vm_state) / 4 */
#define ZKC_VM_WITNESS_WORDS 44  /* sizeof(zkc_vm_cycle_witness) / 4 */

/* answers of the WitnessOracle (main_vm/witness_oracle.rs:45-91) one cycle consumes, flattened by the host before
 * the call.  Only one opcode executes per cycle, so the opcode-specific answers share fields. */
typedef struct zkc_vm_cycle_witness {
    uint32_t code_word[8];       /* get_memory_witness_for_read of the opcode fetch (main_vm/utils.rs:158-181) */
    uint32_t src0_is_pointer;    /* get_memory_witness_for_read of the src0 operand (main_vm/utils.rs:416-440) */
    uint32_t src0_value[8];
    uint32_t callstack_index;    /* ret: which zkc_vm_callstack_witness answers get_callstack_witness (ret.rs:118-160) */
    uint32_t refund;             /* log: get_refunds (log.rs:232-252) */
    uint32_t suggested_page;     /* far_call: get_decommittment_request_suggested_page (far_call.rs:1484-1504) */
    uint32_t value_a[8];         /* uma: get_memory_witness_for_read of cell A (uma.rs:277-312);
                                    log: get_storage_read_witness (log.rs:300-323); far_call: the code hash read (far_call.rs:1203-1223) */
    uint32_t value_b[8];         /* uma: cell B (uma.rs:315-356) */
    uint64_t rollback[4];        /* near_call / far_call: get_rollback_queue_tail_witness_for_call (near_call.rs:70-91, far_call.rs:830-848);
                                    log: get_rollback_queue_witness (log.rs:351-371) */
} zkc_vm_cycle_witness;

/* get_callstack_witness (witness_oracle.rs:80-84): the ExecutionContextRecord a ret pops and the callstack sponge
 * state below it.  Out of line (rets are rare): [n_callstack_witness] per instance, indexed by
 * zkc_vm_cycle_witness.callstack_index.  The forward-log fields of `context` are not part of the record (ignored). */
typedef struct zkc_vm_callstack_witness {
    zkc_vm_context context;
    uint64_t previous_sponge_state[12];
} zkc_vm_callstack_witness;

/* ClosedFormInputWitness<F, VmLocalState, VmInputData, VmOutputData>, fsm_input_output/circuit_inputs/main_vm.rs:9-71 */
typedef struct zkc_vm_closed_form {
    uint32_t start_flag;
    uint32_t completion_flag; /* out */
    /* VmInputData */
    uint64_t rollback_queue_tail_for_block[4];
    uint64_t memory_queue_initial_tail[12];
    uint32_t memory_queue_initial_length, _pad0;
    uint64_t decommitment_queue_initial_tail[12];
    uint32_t decommitment_queue_initial_length;
    uint32_t zkporter_is_available;       /* GlobalContext, vm_state/mod.rs:159-162 */
    uint32_t default_aa_code_hash[8];
    /* VmOutputData (out; expected value if compare_expected) */
    zkc_queue_state4 log_queue_final_state;
    zkc_queue_state12 memory_queue_final_state;
    zkc_queue_state12 decommitment_queue_final_state;
    zkc_vm_state hidden_fsm_input;
    zkc_vm_state hidden_fsm_output;
} zkc_vm_closed_form;

/* trace columns of one vm_cycle (main_vm/cycle.rs:28-795 and pre_state.rs:71-519): the values of the path the
 * cycle takes (the reference evaluates every opcode gadget obliviously and selects; the selected values are the
 * ones that reach the next state).  The nine Poseidon2 relations of a cycle (1 opcode fetch + MAX_SPONGES_PER_CYCLE = 8,
 * state_diffs.rs:15, cycle.rs:937-957) are columns too: slot k holds the permutation OUTPUT when the relation is
 * enforced, zeros otherwise.  Slot use:
 *   0 opcode fetch | 1 src0 read, uma read A, log round 0, callstack round 0 | 2 dst0 write, uma read B, log round 1,
 *   callstack round 1 | 3 uma write A, log round 2 (forward), callstack round 2 | 4 uma write B, log round 2 (rollback),
 *   callstack round 3 | 5-7 far call code-hash read (forward log queue) | 8 far call decommitment queue push.
 * OP_AUX is one block shared by the opcode families (zero for every other opcode):
 *   uma       +0 absolute address, +1 cell index, +2 unalignment, +3 page, +4 skip memory access, +5 set panic,
 *             +6 growth cost, +7 incremented offset, +8 read A (8), +16 read B (8), +24 written A (8), +32 written B (8)
 *   log       +0 packed forward encoding (20), +20 read value (8), +28 execute, +29 execute rollback, +30 ergs to burn
 *   near_call / far_call / ret   +0 new ExecutionContextRecord (42, flatten order), +42 apply near call, +43 apply ret,
 *             +44 ret is panic (after the non-local-frame exceptions), +45 perform revert, +46 apply far call,
 *             +47 far call exception (becomes the pending exception) */
enum zkc_vm_col {
    ZKC_VM_SHOULD_SKIP_CYCLE = 0,    /* pre_state.rs:91-92 */
    ZKC_VM_PENDING_EXCEPTION_IN = 1,
    ZKC_VM_SHOULD_READ_OPCODE = 2,   /* :130-131 */
    ZKC_VM_SUPER_PC = 3,
    ZKC_VM_SUB_PC = 4,
    ZKC_VM_CODE_WORD = 5,            /* 8: after the select with previous_code_word, :177-182 */
    ZKC_VM_OPCODE = 13,              /* 2: after mask_into_nop / mask_into_panic, :216-221 */
    ZKC_VM_VARIANT = 15,
    ZKC_VM_CONDITION_IDX = 16,
    ZKC_VM_CONDITION = 17,
    ZKC_VM_ERGS_COST = 18,
    ZKC_VM_OUT_OF_ERGS = 19,
    ZKC_VM_KERNEL_MODE_EXCEPTION = 20,
    ZKC_VM_STATIC_EXCEPTION = 21,
    ZKC_VM_CALLSTACK_IS_FULL = 22,
    ZKC_VM_EXPLICIT_PANIC = 23,
    ZKC_VM_MASK_INTO_PANIC = 24,
    ZKC_VM_MASK_INTO_NOP = 25,
    ZKC_VM_PROPS = 26,               /* the 48-bit property bit spread after masking, as one integer */
    ZKC_VM_DIRTY_ERGS_LEFT = 27,
    ZKC_VM_SRC0_REG = 28, ZKC_VM_SRC1_REG = 29, ZKC_VM_DST0_REG = 30, ZKC_VM_DST1_REG = 31, /* 4-bit encodings after masking */
    ZKC_VM_IMM0 = 32, ZKC_VM_IMM1 = 33,
    ZKC_VM_SRC0_PAGE = 34, ZKC_VM_SRC0_INDEX = 35, ZKC_VM_SHOULD_READ_SRC0 = 36, ZKC_VM_SP_AFTER_SRC0 = 37,
    ZKC_VM_DST0_PAGE = 38, ZKC_VM_DST0_INDEX = 39, ZKC_VM_DST0_PERFORMS_MEMORY_ACCESS = 40, ZKC_VM_NEW_SP = 41,
    ZKC_VM_SRC0_FROM_MEMORY = 42,    /* 9: is_pointer, value */
    ZKC_VM_SWAP_OPERANDS = 51,
    ZKC_VM_SRC0 = 52,                /* 9: after swap and fat-pointer erasure, :418-472 */
    ZKC_VM_SRC1 = 61,                /* 9 */
    ZKC_VM_DST0 = 70,                /* 9: cycle.rs:200-230 */
    ZKC_VM_DST1 = 79,                /* 9 */
    ZKC_VM_PERFORM_DST0_MEMORY_WRITE = 88, /* :248-254 */
    ZKC_VM_DST0_UPDATE_REGISTER = 89,      /* :312 */
    ZKC_VM_DST1_UPDATE_REGISTER = 90,      /* a gadget flagged a dst1 candidate (should_update_dst1, :177-187); the register WRITE is gated by
                                              the dst1 selector alone (:330): an encoded dst1 register takes ZKC_VM_DST1 (zero when unflagged) */
    ZKC_VM_FLAGS_OUT = 91,           /* 3 */
    ZKC_VM_PENDING_EXCEPTION_OUT = 94,
    ZKC_VM_PC_OUT = 95,
    ZKC_VM_ERGS_OUT = 96,
    ZKC_VM_HEAP_BOUND_OUT = 97, ZKC_VM_AUX_HEAP_BOUND_OUT = 98,   /* cycle.rs:489-520 */
    ZKC_VM_MEMQ_LENGTH_OUT = 99,
    ZKC_VM_DEPTH_OUT = 100,          /* callstack depth after the cycle */
    ZKC_VM_FORWARD_TAIL_OUT = 101,   /* 4 + length: forward log queue after the cycle, cycle.rs:556-577 */
    ZKC_VM_ROLLBACK_HEAD_OUT = 106,  /* 4 + segment length: rollback queue head of the current frame, cycle.rs:580-607 */
    ZKC_VM_SPONGE_ENFORCE = 111,     /* 9 */
    ZKC_VM_SPONGE_FINAL = 120,       /* 9 x 12 */
    ZKC_VM_OP_AUX = 228,             /* 48 */
    ZKC_VM_NUM_COLS = 276
};
#define ZKC_VM_NUM_SPONGES 9
#define ZKC_VM_OP_AUX_COLS 48

#define ZKC_VM_CHK_INVALID_OPCODE (1u << 0)      /* pre_state.rs:291-299 */
#define ZKC_VM_CHK_UNSUPPORTED_OPCODE (1u << 1)  /* a limit of the out-of-circuit run's memory model (zkc_main_vm_simulate) */
#define ZKC_VM_CHK_SNAPSHOT (1u << 2)            /* the host-supplied per-cycle VmLocalState is not what the previous cycle produces */
#define ZKC_VM_CHK_DIV_RELATION (1u << 3)        /* mul_div.rs:321-324 (never fails on computed witnesses) */
#define ZKC_VM_CHK_BOOTLOADER_EXIT (1u << 4)     /* main_vm/mod.rs:119-122 */
#define ZKC_VM_CHK_ROLLBACK_QUEUE (1u << 5)      /* ret.rs:373-404 (head / tail joins), log.rs:627-632 (claimed rollback head) */
#define ZKC_VM_CHK_CALLSTACK (1u << 6)           /* call_ret.rs:279-284 (popped frame does not hash to the stack state), :300 (depth
                                                    underflow), ret.rs:310-311 (ergs overflow), bad callstack_index */
#define ZKC_VM_CHK_LOG_REFUND (1u << 7)          /* log.rs:256 (refund above the pubdata bytes of an initial write) */
#define ZKC_ERR_UNSUPPORTED 7
#define ZKC_ERR_SNAPSHOT_MISMATCH 8

/* Trace layouts.  DENSE: [ZKC_VM_NUM_COLS][limit].  COMPACT: the 117 sponge columns are zero except where a relation is
 * enforced (~1 of 9 slots per cycle), so they travel as a list instead: `trace` is [ZKC_VM_COMPACT_COLS][limit] -- the dense
 * columns below ZKC_VM_SPONGE_ENFORCE keep their index, the OP_AUX block moves down to ZKC_VM_COMPACT_OP_AUX -- and every
 * enforced relation is one zkc_vm_sponge_record in options->sponge_records (host or device memory like `trace`; in no
 * particular order; their number is returned in statuses[0].reserved; ZKC_ERR_INVALID_ARGUMENT-free overflow: records
 * beyond the capacity are dropped and the count still reports how many there were).  Same values, 38 % fewer bytes. */
#define ZKC_VM_TRACE_DENSE 0
#define ZKC_VM_TRACE_COMPACT 1
#define ZKC_VM_COMPACT_COLS (ZKC_VM_NUM_COLS - 9 - 9 * 12)
#define ZKC_VM_COMPACT_OP_AUX (ZKC_VM_OP_AUX - 9 - 9 * 12)
typedef struct zkc_vm_sponge_record {
    uint32_t row;   /* instance * limit + cycle */
    uint32_t slot;  /* 0..8: column block ZKC_VM_SPONGE_FINAL + 12 * slot (and ZKC_VM_SPONGE_ENFORCE + slot = 1) */
    uint64_t out[12];
} zkc_vm_sponge_record;

typedef struct zkc_vm_options {
    uint32_t compare_expected;
    uint32_t trace_layout;                 /* ZKC_VM_TRACE_DENSE / ZKC_VM_TRACE_COMPACT */
    uint64_t sponge_records_capacity;      /* COMPACT: records the buffer below holds */
    zkc_vm_sponge_record *sponge_records;  /* COMPACT: out */
} zkc_vm_options;

/* main_vm_entry_point, main_vm/mod.rs:47-232.
 *   snapshots : [limit + 1] VmLocalState before every cycle and after the last one (what makes one instance data
 *               parallel).  snapshots[0] must be the start state the circuit selects (:85-97),
 *               every snapshots[i + 1] is verified against the cycle's own result.
 *   witness   : [limit] oracle answers
 *   callstack_witness : [n_callstack_witness] frames popped by the rets of the instance (may be NULL when 0)
 *   trace     : column-major [ZKC_VM_NUM_COLS][limit] or NULL */
int zkc_main_vm_entry_point(zkc_ctx *ctx, zkc_vm_closed_form *io, const zkc_vm_isa *isa, const zkc_vm_state *snapshots,
                            const zkc_vm_cycle_witness *witness, const zkc_vm_callstack_witness *callstack_witness,
                            size_t n_callstack_witness, size_t limit, const zkc_vm_options *options, int on_device,
                            uint64_t *trace, uint64_t commitment[ZKC_COMMITMENT_LEN], zkc_status *status);

/* The same over a batch of `n_instances` independent instances of equal `limit` in ONE set of launches (instances
 * only communicate through their closed-form inputs): ios[n], snapshots [n][limit + 1],
 * witness [n][limit], callstack_witness [n][n_callstack_witness] (equal capacity per instance), trace
 * [n][ZKC_VM_NUM_COLS][limit] or NULL, commitments [n][4], statuses [n].  Returns the first non-OK status code. */
int zkc_main_vm_entry_point_batch(zkc_ctx *ctx, zkc_vm_closed_form *ios, size_t n_instances, const zkc_vm_isa *isa,
                                  const zkc_vm_state *snapshots, const zkc_vm_cycle_witness *witness,
                                  const zkc_vm_callstack_witness *callstack_witness, size_t n_callstack_witness, size_t limit,
                                  const zkc_vm_options *options, int on_device, uint64_t *trace, uint64_t *commitments,
                                  zkc_status *statuses);

/* COLUMN (struct-of-arrays) form of the per-cycle inputs -- the layout the cycle kernel reads.  The 32-bit word w of
 * zkc_vm_state (w = byte offset / 4; 64-bit members are two consecutive words, low half first) of snapshot i of instance k is
 * state_words[w * state_stride + k * (limit + 1) + i]; word w of zkc_vm_cycle_witness of cycle i of instance k is
 * witness_words[w * witness_stride + k * limit + i].  A warp's 32 consecutive cycles read every word as one 128-byte line
 * and "the next snapshot" is the neighbouring element, which the record form (1 176-byte records) cannot offer; an
 * out-of-circuit run that emits columns directly saves the transposition the record entry points do on the device.
 * Device memory only; strides in elements, >= n_instances * (limit + 1) resp. n_instances * limit (multiples of 32 keep every
 * line aligned). */
typedef struct zkc_vm_columns {
    const uint32_t *state_words;
    size_t state_stride;
    const uint32_t *witness_words;
    size_t witness_stride;
} zkc_vm_columns;

/* main_vm_entry_point (main_vm/mod.rs:47-232) over a batch of instances whose snapshots / oracle answers are columns in HBM;
 * callstack_witness is device memory too.  Everything else as zkc_main_vm_entry_point_batch. */
int zkc_main_vm_entry_point_columns(zkc_ctx *ctx, zkc_vm_closed_form *ios, size_t n_instances, const zkc_vm_isa *isa,
                                    const zkc_vm_columns *columns, const zkc_vm_callstack_witness *callstack_witness,
                                    size_t n_callstack_witness, size_t limit, const zkc_vm_options *options, int trace_on_device,
                                    uint64_t *trace, uint64_t *commitments, zkc_status *statuses);

/* records -> columns on the device (what the record entry points run internally): snapshots [n_instances][limit + 1],
 * witness [n_instances][limit], all four pointers device memory.  Asynchronous on the context's stream. */
int zkc_main_vm_rows_to_columns(zkc_ctx *ctx, const zkc_vm_state *snapshots, const zkc_vm_cycle_witness *witness, size_t n_instances,
                                size_t limit, uint32_t *state_words, size_t state_stride, uint32_t *witness_words, size_t witness_stride);

/* ---- transport forms of the main_vm call over PCIe (host buffers) ---------------------------------------------------
 * The record form costs 1 176 + 176 bytes per cycle on the way in and 8 bytes per cell on the way out; both directions are
 * PCIe-bound by a factor of ten against the kernels.  The STREAM forms carry the same values in fewer bytes:
 *
 * IN -- zkc_vm_input_stream: one instance's snapshots / oracle answers as a sequence of SEGMENTS of `segment_cycles` cycles
 * (the last one may be shorter); segment s covers cycles [s * segment_cycles, ...) and the snapshots before each of them
 * plus the one behind the last (a boundary snapshot is in both neighbours).  An out-of-circuit run emits a segment every
 * `segment_cycles` cycles; the call copies, expands and evaluates them as they come (segment = pipeline chunk).
 * A segment is ONE contiguous blob (one asynchronous copy): the columns of zkc_vm_columns restricted to the segment, each
 * 32-bit word either DENSE (all its values) or SPARSE.  A sparse STATE word is a list of (local snapshot index, value)
 * changes: record j of word w holds from index[j] up to the next record of w (or the end of the segment); the first record of
 * a word is at local index 0.  A sparse WITNESS word is the list of (local cycle index, value) of its non-zero values.
 * Blob layout (all offsets in bytes from the start of the blob, every array 16-byte aligned):
 *   zkc_vm_segment_header
 *   uint16_t dense_state_word[n_dense_state]        increasing
 *   uint16_t dense_witness_word[n_dense_witness]
 *   uint32_t dense_state[n_dense_state][n_cycles + 1]
 *   uint32_t dense_witness[n_dense_witness][n_cycles]
 *   uint32_t sparse_state_offsets[ZKC_VM_STATE_WORDS + 1]     records [offsets[w], offsets[w + 1]) belong to word w
 *   uint32_t sparse_state_index[n_sparse_state], sparse_state_value[n_sparse_state]
 *   uint32_t sparse_witness_offsets[ZKC_VM_WITNESS_WORDS + 1]
 *   uint32_t sparse_witness_index[n_sparse_witness], sparse_witness_value[n_sparse_witness]
 * zkc_vm_encode_input_stream derives the stream from records (a word goes dense when that is fewer bytes: 4 per value
 * against 8 per change). */
#define ZKC_VM_SEGMENT_MAGIC 0x5a4b5347u
typedef struct zkc_vm_segment_header {
    uint32_t magic, first_cycle, n_cycles, n_dense_state, n_dense_witness, n_sparse_state, n_sparse_witness, reserved;
    uint32_t off_dense_state_word, off_dense_witness_word, off_dense_state, off_dense_witness;
    uint32_t off_sparse_state_offsets, off_sparse_state_index, off_sparse_state_value;
    uint32_t off_sparse_witness_offsets, off_sparse_witness_index, off_sparse_witness_value;
    uint32_t blob_bytes, reserved2;
} zkc_vm_segment_header;
typedef struct zkc_vm_input_segment {
    const void *blob;  /* starts with a zkc_vm_segment_header; pinned host memory makes the copy asynchronous */
    uint64_t blob_bytes;
} zkc_vm_input_segment;
typedef struct zkc_vm_input_stream {
    uint64_t limit;
    uint32_t segment_cycles, n_segments;
    const zkc_vm_input_segment *segments;
} zkc_vm_input_stream;

/* records of ONE instance (host memory: snapshots [limit + 1], witness [limit]) -> a stream in pinned host memory owned by
 * the library; free with zkc_vm_input_stream_free.  segment_cycles = 0 picks the default (2^16).  bytes_out (may be NULL):
 * what crosses PCIe. */
int zkc_vm_encode_input_stream(const zkc_vm_state *snapshots, const zkc_vm_cycle_witness *witness, size_t limit, size_t segment_cycles,
                               zkc_vm_input_stream **out, uint64_t *bytes_out);
void zkc_vm_input_stream_free(zkc_vm_input_stream *stream);

/* OUT -- the PACKED trace: the same cells as the DENSE layout, every column in the narrowest unsigned type its values fit
 * (booleans / small integers u8, 16-bit values u16, limbs u32, field elements u64), column-major per type block:
 * cols8[slot * rows + g] with g = instance * limit + row, rows = n_instances * limit.  zkc_vm_packed_layout gives
 * (kind, slot) of every DENSE-layout column.  The opcode-family block (ZKC_VM_OP_AUX, zero on ~85 % of the rows) and the
 * forward / rollback queue ends (which move only on log / call / ret rows) travel as one zkc_vm_aux_record per row that has
 * either: a row without a record has a zero OP_AUX block and the queue ends of the row before it (row 0 of an instance always
 * has a record).  The sponge columns travel as zkc_vm_sponge_record like in the COMPACT layout.  Records arrive in no
 * particular order; the counts are written to n_aux_records / n_sponge_records (records beyond a capacity are dropped, the
 * count still says how many there were). */
#define ZKC_VM_TRACE_PACKED 2
enum zkc_vm_packed_kind { ZKC_VM_PK_U8 = 0, ZKC_VM_PK_U16, ZKC_VM_PK_U32, ZKC_VM_PK_U64, ZKC_VM_PK_AUX_RECORD, ZKC_VM_PK_SPONGE_RECORD,
                          ZKC_VM_PK_LIMB_RECORD };
/* Three 9-column groups of u32 limbs that are constant or zero on most rows travel as one zkc_vm_limb_record per row that
 * changes them: the code word (rows without a record carry the code word of the row before; a row that fetches its opcode and
 * row 0 of an instance always have one), the src0 memory operand and dst1 (rows without a record hold zeros). */
#define ZKC_VM_LIMB_CODE_WORD 0        /* v[0..8) = ZKC_VM_CODE_WORD .. + 7, v[8] = 0 */
#define ZKC_VM_LIMB_SRC0_FROM_MEMORY 1 /* v[0..9) = ZKC_VM_SRC0_FROM_MEMORY .. + 8 (is_pointer, 8 limbs) */
#define ZKC_VM_LIMB_DST1 2             /* v[0..9) = ZKC_VM_DST1 .. + 8 */
typedef struct zkc_vm_limb_record {
    uint32_t row;  /* instance * limit + cycle */
    uint32_t kind; /* ZKC_VM_LIMB_* */
    uint32_t v[9];
    uint32_t reserved;
} zkc_vm_limb_record;
typedef struct zkc_vm_aux_record {
    uint32_t row;      /* instance * limit + cycle */
    uint32_t reserved;
    uint64_t op_aux[48];      /* ZKC_VM_OP_AUX .. + 47 */
    uint64_t queue_ends[10];  /* ZKC_VM_FORWARD_TAIL_OUT .. + 4, ZKC_VM_ROLLBACK_HEAD_OUT .. + 4 */
} zkc_vm_aux_record;
typedef struct zkc_vm_packed_trace {
    uint8_t *cols8;
    uint16_t *cols16;
    uint32_t *cols32;
    uint64_t *cols64;
    zkc_vm_aux_record *aux_records;
    uint64_t aux_capacity, n_aux_records;        /* in, out */
    zkc_vm_sponge_record *sponge_records;
    uint64_t sponge_capacity, n_sponge_records;  /* in, out */
    zkc_vm_limb_record *limb_records;
    uint64_t limb_capacity, n_limb_records;      /* in, out */
} zkc_vm_packed_trace;
/* kind[c], slot[c] for c < ZKC_VM_NUM_COLS; counts[k] = columns of kind k (k < 4: the heights of the four type blocks; a
 * LIMB_RECORD column's slot = 9 * ZKC_VM_LIMB_* kind + its index in v[]) */
void zkc_vm_packed_layout(uint8_t kind[ZKC_VM_NUM_COLS], uint16_t slot[ZKC_VM_NUM_COLS], uint32_t counts[7]);

/* main_vm_entry_point (main_vm/mod.rs:47-232) over host buffers in the stream forms: streams[n_instances] (equal limit and
 * segment_cycles), callstack_witness [n_instances][n_callstack_witness] host records, `out` (may be NULL: no witness
 * wanted) host buffers.  Segment by segment, H2D | expansion + kernels + packing | D2H overlap; everything else as
 * zkc_main_vm_entry_point_batch. */
int zkc_main_vm_entry_point_stream(zkc_ctx *ctx, zkc_vm_closed_form *ios, size_t n_instances, const zkc_vm_isa *isa,
                                   const zkc_vm_input_stream *const *streams, const zkc_vm_callstack_witness *callstack_witness,
                                   size_t n_callstack_witness, size_t limit, const zkc_vm_options *options, zkc_vm_packed_trace *out,
                                   uint64_t *commitments, zkc_status *statuses);

/* Constraint evaluation of finished main_vm traces (DENSE layout, [n_instances][ZKC_VM_NUM_COLS][limit]): every relation
 * that is local to a row -- booleanity / ranges of the allocated cells, opcode decoding against the ISA tables
 * (decoded_opcode.rs:395-527) and the exception masks (:120-157), AddSubRelation / MulDivRelation / bitwise relations of the
 * arithmetic opcodes (opcodes/mod.rs:101-180, binop.rs), the dot-product selection of dst0 / dst1 (cycle.rs:199-246), zero
 * sponge columns where nothing is enforced.  violations = number of rows with at least one failing relation;
 * status->failed_checks = which families (ZKC_VMV_*), status->first_bad_row = instance * limit + row.  A pure stream over
 * the trace: 2 208 bytes per row, HBM-bound. */
#define ZKC_VMV_BOOLEAN (1u << 0)
#define ZKC_VMV_RANGE (1u << 1)
#define ZKC_VMV_DECODE (1u << 2)
#define ZKC_VMV_EXCEPTION_MASKS (1u << 3)
#define ZKC_VMV_ADD_SUB (1u << 4)
#define ZKC_VMV_MUL_DIV (1u << 5)
#define ZKC_VMV_BINOP (1u << 6)
#define ZKC_VMV_FLAGS (1u << 7)
#define ZKC_VMV_SELECTION (1u << 8)
#define ZKC_VMV_SPONGE (1u << 9)
int zkc_main_vm_check_trace(zkc_ctx *ctx, const zkc_vm_isa *isa, const uint64_t *trace, size_t limit, size_t n_instances,
                            int on_device, uint64_t *violations, zkc_status *status);

/* ---- cells of the arithmetic opcode gadgets, evaluated OBLIVIOUSLY --------------------------------------------------------
 * The reference runs every opcode gadget on every cycle and selects afterwards (main_vm/cycle.rs:73-156); the DENSE trace
 * (zkc_vm_col) carries the selected path.  zkc_main_vm_gadget_cells produces, for every cycle, the cells the add/sub, binop,
 * mul/div and shift gadgets allocate regardless of the opcode, and the relations vm_cycle enforces once per cycle:
 *   register_input_view.rs:27-53   byte views of src0 / src1
 *   opcodes/add_sub.rs:8-166       both results, the selected result, the shuffled AddSubRelation, flags
 *   opcodes/binop.rs:14-244        the 32 composite lookups, their 96-cell decomposition, and / or / xor limbs, result, flags
 *   opcodes/mul_div.rs:199-417     product, quotient / remainder (divisor 0: quotient 0, remainder = src0, :119-123), selections,
 *                                  the MulDivRelation and the remainder < divisor AddSubRelation, flags
 *   opcodes/shifts.rs:8-198        shift constant (tables/bitshift.rs), both shift directions, relations, result, flags
 *   cycle.rs:619-670               the conditional range check, the ONE enforced AddSubRelation (candidates add_sub, mul_div,
 *                                  shifts; the last pushed is the default, opcodes/mod.rs:101-125: 8 carries) and the ONE enforced
 *                                  MulDivRelation (candidates mul_div, shifts; opcodes/mod.rs:129-180: 64 (low, high) partial
 *                                  products of UInt32::fma_with_carry + 8 row-end sums)
 * Inputs are the src0 / src1 operand and property-bit columns of the DENSE trace (ZKC_VM_SRC0, ZKC_VM_SRC1, ZKC_VM_PROPS).
 * The ptr, jump and context gadgets' cells are the second block (zkc_main_vm_state_gadget_cells below); the UMA, log and
 * call/ret gadgets' non-selected cells are not produced.
 * X(name, width): column ZKC_VMG_<name> .. + width - 1 of the gadget block [ZKC_VMG_NUM_COLS][limit]. */
#define ZKC_VM_GADGET_COLUMNS(X) \
    X(SRC0_BYTES, 32) X(SRC1_BYTES, 32) \
    X(ADD_RESULT, 8) X(ADD_OF, 1) X(SUB_RESULT, 8) X(SUB_UF, 1) X(ADDSUB_RESULT, 8) X(ADDSUB_NEW_B, 8) X(ADDSUB_NEW_C, 8) \
    X(ADDSUB_NEW_OF, 1) X(ADDSUB_LIMB_IS_ZERO, 8) X(ADDSUB_RESULT_IS_ZERO, 1) X(ADDSUB_GT, 1) X(ADDSUB_APPLY_ANY, 1) \
    X(ADDSUB_UPDATE_FLAGS, 1) \
    X(BINOP_COMPOSITE, 32) X(BINOP_ALL_RESULTS, 96) X(BINOP_AND, 8) X(BINOP_OR, 8) X(BINOP_XOR, 8) X(BINOP_RESULT, 8) \
    X(BINOP_LIMB_IS_ZERO, 8) X(BINOP_RESULT_IS_ZERO, 1) X(BINOP_UPDATE_FLAGS, 1) \
    X(MUL_LOW, 8) X(MUL_HIGH, 8) X(DIV_QUOTIENT, 8) X(DIV_REMAINDER, 8) X(MULDIV_RESULT_0, 8) X(MULDIV_RESULT_1_UNMASKED, 8) \
    X(MULDIV_REM_TO_ENFORCE, 8) X(MULDIV_A_TO_ENFORCE, 8) X(MULDIV_MUL_LOW_TO_ENFORCE, 8) X(MULDIV_MUL_HIGH_TO_ENFORCE, 8) \
    X(MUL_HIGH_IS_ZERO, 1) X(MUL_LOW_IS_ZERO, 1) X(MUL_OF, 1) X(MUL_GT, 1) X(DIV_DIVISOR_IS_ZERO, 1) X(DIV_QUOTIENT_IS_ZERO, 1) \
    X(DIV_REMAINDER_IS_ZERO, 1) X(DIV_SUB_RESULT, 8) X(DIV_REMAINDER_IS_LESS, 1) X(DIV_MASK_REMAINDER, 1) X(MULDIV_RESULT_1, 8) \
    X(DIV_EQ, 1) X(DIV_GT, 1) X(MULDIV_OF, 1) X(MULDIV_EQ, 1) X(MULDIV_GT, 1) X(MULDIV_APPLY_ANY, 1) X(MULDIV_SET_FLAGS, 1) \
    X(SHIFT_AMOUNT, 1) X(SHIFT_IS_ZERO, 1) X(SHIFT_INVERTED, 1) X(SHIFT_CHANGE_FLAG, 1) X(SHIFT_FULL, 1) X(SHIFT_CONSTANT, 8) \
    X(SHIFT_IS_RIGHT, 1) X(SHIFT_RSHIFT_Q, 8) X(SHIFT_RSHIFT_R, 8) X(SHIFT_APPLY_LEFT, 1) X(SHIFT_LSHIFT_LOW, 8) \
    X(SHIFT_LSHIFT_HIGH, 8) X(SHIFT_REM_TO_ENFORCE, 8) X(SHIFT_A_TO_ENFORCE, 8) X(SHIFT_MUL_LOW_TO_ENFORCE, 8) \
    X(SHIFT_MUL_HIGH_TO_ENFORCE, 8) X(SHIFT_SUB_RESULT, 8) X(SHIFT_REMAINDER_IS_LESS, 1) X(SHIFT_TEMP_RESULT, 8) \
    X(SHIFT_RESULT, 8) X(SHIFT_RESULT_IS_ZERO, 1) X(SHIFT_SET_FLAGS, 1) \
    X(RANGE_CHECK, 8) X(ADDREL_A, 8) X(ADDREL_B, 8) X(ADDREL_C, 8) X(ADDREL_OF, 1) X(ADDREL_CARRY, 8) \
    X(MULREL_A, 8) X(MULREL_B, 8) X(MULREL_REM, 8) X(MULREL_LOW, 8) X(MULREL_HIGH, 8) X(MULREL_PARTIAL_LOW, 64) \
    X(MULREL_PARTIAL_HIGH, 64) X(MULREL_ROW_END, 8)
enum zkc_vm_gadget_col {
#define ZKC_VMG_X(name, width) ZKC_VMG_##name, ZKC_VMG_##name##_LAST = ZKC_VMG_##name + (width)-1,
    ZKC_VM_GADGET_COLUMNS(ZKC_VMG_X)
#undef ZKC_VMG_X
    ZKC_VMG_NUM_COLS
};
/* trace: DENSE traces [n_instances][ZKC_VM_NUM_COLS][limit] (host, or device with on_device != 0);
 * gadget_trace: out, [n_instances][ZKC_VMG_NUM_COLS][limit] in the same memory space */
int zkc_main_vm_gadget_cells(zkc_ctx *ctx, const uint64_t *trace, size_t limit, size_t n_instances, int on_device, uint64_t *gadget_trace);

/* ---- cells of the ptr, jump and context opcode gadgets, evaluated OBLIVIOUSLY ------------------------------------------------
 * The second gadget-cell block: what apply_ptr, apply_jump and apply_context allocate on EVERY cycle whatever the opcode
 * (apply_nop allocates nothing, opcodes/nop.rs:4-24).  Unlike the arithmetic gadgets these read the VM state the cycle starts
 * from, so the call takes the per-cycle snapshots next to the finished DENSE trace:
 *   opcodes/ptr.rs:6-183      operand type / range checks of src1, the three overflowing add / sub results, the panic conditions,
 *                             the selected limbs and the dst0 candidate (is_pointer of src0 + 8 limbs)
 *   opcodes/jump.rs:3-38      the UInt16 destination recomposed from the two low bytes of src0
 *   opcodes/context.rs:7-307  the three state-update flags, read_only / write_like / write_to_dst0, tx_number_in_block + 1, the
 *                             meta word's highest limb, and the widening select chain low_u32 -> 128 -> 160 (this, caller, code
 *                             address) -> 256 (meta) that ends in the dst0 candidate
 * Inputs: ZKC_VM_SRC0 / ZKC_VM_SRC1 (is_pointer + limbs, after swap and fat-pointer erasure), ZKC_VM_PROPS, ZKC_VM_NEW_SP (the
 * draft state's sp, pre_state.rs:370) and ZKC_VM_DIRTY_ERGS_LEFT (preliminary_ergs_left, pre_state.rs:513) of the DENSE trace;
 * tx_number_in_block, ergs_per_pubdata_byte and the current context's addresses, shard ids, heap bounds and context_u128 of
 * snapshot i (pre_state.rs does not touch them).
 * X(name, width): column ZKC_VMS_<name> .. + width - 1 of the block [ZKC_VMS_NUM_COLS][limit]. */
#define ZKC_VM_STATE_GADGET_COLUMNS(X) \
    X(PTR_SRC1_IS_INTEGER, 1) X(PTR_ARGS_VALID, 1) X(PTR_ARGS_INVALID, 1) X(PTR_SRC1_LIMB_IS_ZERO, 8) X(PTR_SRC1_32_256_IS_ZERO, 1) \
    X(PTR_SRC1_0_128_IS_ZERO, 1) X(PTR_SRC1_32_256_IS_NONZERO, 1) X(PTR_ARITH_VARIANT, 1) X(PTR_TOO_LARGE_OFFSET, 1) \
    X(PTR_SRC1_0_128_IS_NONZERO, 1) X(PTR_DIRTY_PACK, 1) X(PTR_ADD_RESULT, 1) X(PTR_ADD_OF, 1) X(PTR_ADD_PANIC, 1) \
    X(PTR_SUB_RESULT, 1) X(PTR_SUB_UF, 1) X(PTR_SUB_PANIC, 1) X(PTR_SHRINK_RESULT, 1) X(PTR_SHRINK_UF, 1) X(PTR_SHRINK_PANIC, 1) \
    X(PTR_ANY_PANIC, 1) X(PTR_SHOULD_PANIC, 1) X(PTR_OK, 1) X(PTR_UPDATE_REGISTER, 1) X(PTR_LOW_IF_ADD, 1) X(PTR_LOW_IF_ADD_OR_SUB, 1) \
    X(PTR_96_128_IF_SHRINK, 1) X(PTR_HIGHEST_128, 4) X(PTR_LOWEST32, 1) X(PTR_96_128, 1) X(PTR_DST0, 9) \
    X(JUMP_DST, 1) \
    X(CTX_WRITE_TO_CONTEXT, 1) X(CTX_SET_PUBDATA_ERGS, 1) X(CTX_INCREMENT_TX, 1) X(CTX_READ_ONLY, 1) X(CTX_WRITE_LIKE, 1) \
    X(CTX_WRITE_TO_DST0, 1) X(CTX_INCREMENTED_TX_NUMBER, 1) X(CTX_TX_OF, 1) X(CTX_META_HIGHEST, 1) X(CTX_LOW_U32, 1) \
    X(CTX_RESULT_128, 4) X(CTX_RESULT_160_THIS, 5) X(CTX_RESULT_160_CALLER, 5) X(CTX_RESULT_160_CODE, 5) X(CTX_RESULT_256, 8)
enum zkc_vm_state_gadget_col {
#define ZKC_VMS_X(name, width) ZKC_VMS_##name, ZKC_VMS_##name##_LAST = ZKC_VMS_##name + (width)-1,
    ZKC_VM_STATE_GADGET_COLUMNS(ZKC_VMS_X)
#undef ZKC_VMS_X
    ZKC_VMS_NUM_COLS
};
/* trace: DENSE traces [n_instances][ZKC_VM_NUM_COLS][limit]; snapshots: [n_instances][limit + 1] records (the ones
 * zkc_main_vm_entry_point took); both host, or both device with on_device != 0;
 * gadget_trace: out, [n_instances][ZKC_VMS_NUM_COLS][limit] in the same memory space */
int zkc_main_vm_state_gadget_cells(zkc_ctx *ctx, const uint64_t *trace, const zkc_vm_state *snapshots, size_t limit, size_t n_instances,
                                   int on_device, uint64_t *gadget_trace);

/* ---- the memory-queue relations every cycle evaluates, OBLIVIOUSLY -------------------------------------------------------------
 * Three of the nine Poseidon2 relations of a cycle do not belong to an opcode gadget: the opcode fetch, the src0 read and the dst0
 * write.  The reference builds each one on EVERY cycle whatever the access flag -- the query is encoded, absorbed with replacement
 * into the current memory-queue tail (initial_state = encoding || tail[8..12]), the permutation computed, and the new tail / length
 * SELECTED by the flag:
 *   may_be_read_memory_for_code             main_vm/utils.rs:129-233   (R::compute_round_function on every cycle, :212)
 *   may_be_read_memory_for_source_operand   main_vm/utils.rs:388-522   (initial_state :479-492, selects :499-511)
 *   may_be_write_memory                     main_vm/cycle.rs:799-935   (initial_state :871-884, selects :891-903)
 *   enforce_sponges                         main_vm/cycle.rs:937-957   (true_final = R(initial_state) of the candidate selected
 *                                                                       for the slot, cycle.rs:673-721)
 * The DENSE trace (zkc_vm_col) holds a slot's permutation output only when the relation is ENFORCED (ZKC_VM_SPONGE_ENFORCE);
 * this block holds, for every cycle, the three initial states, R(initial_state), the selected tail and length after each step.
 * A memory value that is not read is zero (the oracle's answer for execute = false).  SELECTED = 1 when slots 1 / 2 of
 * enforce_sponges evaluate these two candidates, i.e. no opcode with its own sponges applies (UMA, log, near_call, far_call, ret:
 * cycle.rs:681-721 -- then the slot's in-circuit permutation runs on the opcode's candidate and SRC0_FINAL / DST0_FINAL here are
 * not cells of the reference; INIT / STATE_AFTER / LENGTH_AFTER are on every cycle).  The fetch's permutation is a cell on every
 * cycle.  The remaining six slots depend on the non-selected cells of those gadgets (not produced).
 * Inputs: SHOULD_READ_OPCODE, SUPER_PC, CODE_WORD, SRC0_PAGE / INDEX, SHOULD_READ_SRC0, SRC0_FROM_MEMORY, DST0_PAGE / INDEX,
 * PERFORM_DST0_MEMORY_WRITE, DST0, PROPS of the DENSE trace; timestamp, code_page, memory_queue_state / length of snapshot i.
 * X(name, width): column ZKC_VMQ_<name> .. + width - 1 of the block [ZKC_VMQ_NUM_COLS][limit]. */
#define ZKC_VM_MEMORY_SPONGE_COLUMNS(X) \
    X(SELECTED, 1) \
    X(FETCH_INIT, 12) X(FETCH_FINAL, 12) X(FETCH_STATE_AFTER, 12) X(FETCH_LENGTH_AFTER, 1) \
    X(SRC0_INIT, 12) X(SRC0_FINAL, 12) X(SRC0_STATE_AFTER, 12) X(SRC0_LENGTH_AFTER, 1) \
    X(DST0_INIT, 12) X(DST0_FINAL, 12) X(DST0_STATE_AFTER, 12) X(DST0_LENGTH_AFTER, 1)
enum zkc_vm_memory_sponge_col {
#define ZKC_VMQ_X(name, width) ZKC_VMQ_##name, ZKC_VMQ_##name##_LAST = ZKC_VMQ_##name + (width)-1,
    ZKC_VM_MEMORY_SPONGE_COLUMNS(ZKC_VMQ_X)
#undef ZKC_VMQ_X
    ZKC_VMQ_NUM_COLS
};
/* trace: DENSE traces [n_instances][ZKC_VM_NUM_COLS][limit]; snapshots: [n_instances][limit + 1] records; both host, or both
 * device with on_device != 0; sponge_trace: out, [n_instances][ZKC_VMQ_NUM_COLS][limit] in the same memory space */
int zkc_main_vm_memory_sponge_cells(zkc_ctx *ctx, const uint64_t *trace, const zkc_vm_state *snapshots, size_t limit, size_t n_instances,
                                    int on_device, uint64_t *sponge_trace);

/* ---- cells of create_prestate that are not columns of the DENSE trace -------------------------------------------------------------
 * The DENSE trace names the RESULTS of the cycle's preamble (main_vm/pre_state.rs:71-519); this block holds the cells on the way:
 *   pre_state.rs:88-105     execute_cycle, should_try_to_read_opcode, the pending exception taken down
 *   pre_state.rs:107-129    pc + 1; should_read_memory (main_vm/utils.rs:106-120): page / super-pc equalities, can_skip, its negation
 *   pre_state.rs:143-156    the four timestamps after the cycle's first, and the selected next-cycle timestamp
 *   pre_state.rs:183-214    the three sub-pc mask bits and the three-stage select of the 64-bit opcode inside the code word
 *                           (the last stage is the opcode BEFORE mask_into_nop / mask_into_panic)
 *   decoded_opcode.rs:192-202  the four 15-bit register selector masks (reg_idx_into_bitspread + spread_into_bits)
 *   pre_state.rs:303-329    the two 15-step VMRegister::conditionally_select chains that pick src0 / src1 out of the register
 *                           file (15 x 9 cells each), the 15-step chain for the low limb of the dst0 register, both low_u16
 *   pre_state.rs:337-343    stack / heap / aux heap page
 *   main_vm/utils.rs:237-305   source location: absolute_mode, index_for_absolute, index_for_relative, use_stack, did_read before
 *                           the NOP rule, not_nop
 *   main_vm/utils.rs:307-386   destination location: index_for_absolute, index_for_relative_with_push, index_for_relative,
 *                           did_write before the NOP rule, the push-or-relative index
 *   pre_state.rs:403-413    src0 after the use_reg select and after the use_imm select
 *   pre_state.rs:418-455    is_assymmetric, t0, t1 of swap_operands; both operands after the swap
 *   pre_state.rs:457-479    not_kernel_mode, the pointer-keeping opcode class, should_erase, the two erase flags (applying them to
 *                           the swapped operands gives ZKC_VM_SRC0 / ZKC_VM_SRC1 of the DENSE trace)
 * Inputs: the DENSE trace (skip / pending flags, sub-pc, code word, register indices, immediates, property bits, the src0 memory
 * value, swap flag) and snapshot i (registers, pc, sp, pages, previous code page / super-pc, timestamp, kernel mode).
 * X(name, width): column ZKC_VMP_<name> .. + width - 1 of the block [ZKC_VMP_NUM_COLS][limit]. */
#define ZKC_VM_PRESTATE_COLUMNS(X) \
    X(EXECUTE_CYCLE, 1) X(SHOULD_TRY_TO_READ_OPCODE, 1) X(PENDING_EXCEPTION_TAKEN_DOWN, 1) X(PC_PLUS_ONE, 1) X(PC_PLUS_ONE_OF, 1) \
    X(CODE_PAGES_ARE_EQUAL, 1) X(SUPER_PC_ARE_EQUAL, 1) X(CAN_SKIP_READ, 1) X(SHOULD_READ_FOR_NEW_PC, 1) \
    X(TIMESTAMPS, 4) X(NEXT_CYCLE_TIMESTAMP, 1) X(SUBPC_BITMASK, 3) X(OPCODE_SELECT_CHAIN, 6) \
    X(SRC0_SELECTORS, 15) X(SRC1_SELECTORS, 15) X(DST0_SELECTORS, 15) X(DST1_SELECTORS, 15) \
    X(DRAFT_SRC0_CHAIN, 135) X(SRC1_REGISTER_CHAIN, 135) X(DST0_REG_LOW_CHAIN, 15) X(SRC0_REG_LOWEST, 1) X(DST0_REG_LOWEST, 1) \
    X(STACK_PAGE, 1) X(HEAP_PAGE, 1) X(AUX_HEAP_PAGE, 1) \
    X(SRC_ABSOLUTE_MODE, 1) X(SRC_INDEX_FOR_ABSOLUTE, 1) X(SRC_INDEX_FOR_RELATIVE, 1) X(SRC_USE_STACK, 1) X(SRC_DID_READ_UNMASKED, 1) X(NOT_NOP, 1) \
    X(DST_INDEX_FOR_ABSOLUTE, 1) X(DST_INDEX_FOR_RELATIVE_WITH_PUSH, 1) X(DST_INDEX_FOR_RELATIVE, 1) X(DST_DID_WRITE_UNMASKED, 1) \
    X(DST_INDEX_SOMEWHAT_RELATIVE, 1) \
    X(SRC0_AFTER_USE_REG, 9) X(SRC0_AFTER_USE_IMM, 9) X(SWAP_IS_ASSYMMETRIC, 1) X(SWAP_T0, 1) X(SWAP_T1, 1) X(SRC0_SWAPPED, 9) X(SRC1_SWAPPED, 9) \
    X(NOT_KERNEL_MODE, 1) X(KEEPS_POINTERS, 1) X(SHOULD_ERASE, 1) X(SHOULD_ERASE_SRC0, 1) X(SHOULD_ERASE_SRC1, 1)
enum zkc_vm_prestate_col {
#define ZKC_VMP_X(name, width) ZKC_VMP_##name, ZKC_VMP_##name##_LAST = ZKC_VMP_##name + (width)-1,
    ZKC_VM_PRESTATE_COLUMNS(ZKC_VMP_X)
#undef ZKC_VMP_X
    ZKC_VMP_NUM_COLS
};
/* trace: DENSE traces [n_instances][ZKC_VM_NUM_COLS][limit]; snapshots: [n_instances][limit + 1] records; both host, or both
 * device with on_device != 0; prestate_trace: out, [n_instances][ZKC_VMP_NUM_COLS][limit] in the same memory space */
int zkc_main_vm_prestate_cells(zkc_ctx *ctx, const uint64_t *trace, const zkc_vm_state *snapshots, size_t limit, size_t n_instances,
                               int on_device, uint64_t *prestate_trace);

/* ---- the register write-back of the state diffs (cycle.rs:158-433) -----------------------------------------------------------------
 * What vm_cycle allocates when it applies dst0 / dst1 and the far call / far return register conventions to the 15 registers:
 *   cycle.rs:160-189   dst0_update_potentially_to_memory, can_update_dst0_as_register_only (the two multi_or over the candidates' flags)
 *   cycle.rs:298-304   dst0_performs_reg_update, t (ZKC_VM_DST0_UPDATE_REGISTER of the DENSE trace = can_update_dst0_as_register_only | t)
 *   cycle.rs:323-330   per register: write_as_dst0 = dst0_update_register & selector (write_as_dst1 IS the dst1 selector bit,
 *                      ZKC_VMP_DST1_SELECTORS of the prestate block: an encoded dst1 register is written whatever the gadgets flagged)
 *   cycle.rs:349-379   the specific updates of a far call (r1, r2: far_call.rs:1041-1043) and a far return (r1: ret.rs:444-445), the
 *                      remove-pointer marker and the zero-out flag: multi_or over the far call's and the far return's requests
 *                      (far_call.rs:1045-1070, ret.rs:451-462)
 *   cycle.rs:381-421   any_ptr_update_as_dst0, the is_pointer dot product over the dst0-side candidates, is_pointer after that select,
 *                      the dst1 dot product, is_pointer after the dst1 select (= the register's marker in the next state)
 *   cycle.rs:423-432   the value select chain: after dst0, after the far call's update (r1, r2 only), after the far return's update
 *                      (r1 only), after the zero-out, after dst1 (= the register's value in the next state)
 * Inputs: the DENSE trace (property bits, dst0 / dst1 dot products, their register indices, the update / memory-access flags, the
 * far call's ABI word in src0 and target in src1), snapshot i (registers, is_local_call, is_kernel_mode), the calling-convention
 * register lists of the ISA tables, and snapshot i + 1 for ONE thing: the far call's / far return's new r1 when that update is
 * applied (final_fat_ptr.into_register, far_call.rs:1008, ret.rs:441: a cell of the call / ret gadget, which this block does not
 * re-derive; the entry point's snapshot link has verified it).  The reference allocates no zero-out select for a register no list
 * names and no far-call step for r3..r15; the block keeps one uniform layout and passes the value through there.
 * X(name, width): column ZKC_VMW_<name> .. + width - 1 of the block [ZKC_VMW_NUM_COLS][limit]. */
#define ZKC_VM_WRITEBACK_COLUMNS(X) \
    X(DST0_UPDATE_POTENTIALLY_TO_MEMORY, 1) X(CAN_UPDATE_DST0_AS_REGISTER_ONLY, 1) X(DST0_PERFORMS_REG_UPDATE, 1) X(DST0_REG_UPDATE_T, 1) \
    X(FAR_CALL_UPDATE, 1) X(FAR_CALL_NON_SYSTEM, 1) X(FAR_CALL_CLEANUP_REGISTER, 1) X(FAR_RETURN_UPDATE, 1) X(FAR_CALL_NEW_R2_LOW, 1) \
    X(WRITE_AS_DST0, 15) X(REMOVE_PTR_MARKER, 15) X(ZERO_OUT, 15) X(ANY_PTR_UPDATE_AS_DST0, 15) X(IS_PTR_AS_DST0, 15) X(IS_PTR_AFTER_DST0, 15) \
    X(IS_PTR_AS_DST1, 15) X(IS_PTR_AFTER_DST1, 15) \
    X(VALUE_AFTER_DST0, 120) X(VALUE_AFTER_FAR_CALL, 16) X(VALUE_AFTER_FAR_RETURN, 8) X(VALUE_AFTER_ZERO_OUT, 120) X(VALUE_AFTER_DST1, 120)
enum zkc_vm_writeback_col {
#define ZKC_VMW_X(name, width) ZKC_VMW_##name, ZKC_VMW_##name##_LAST = ZKC_VMW_##name + (width)-1,
    ZKC_VM_WRITEBACK_COLUMNS(ZKC_VMW_X)
#undef ZKC_VMW_X
    ZKC_VMW_NUM_COLS
};
/* trace / snapshots as for zkc_main_vm_prestate_cells; isa: host pointer (its calling-convention register lists are read on the host);
 * writeback_trace: out, [n_instances][ZKC_VMW_NUM_COLS][limit] in the memory space of trace */
int zkc_main_vm_writeback_cells(zkc_ctx *ctx, const zkc_vm_isa *isa, const uint64_t *trace, const zkc_vm_state *snapshots, size_t limit,
                                size_t n_instances, int on_device, uint64_t *writeback_trace);

/* the state main_vm_entry_point starts from when start_flag is set: initial_bootloader_state, main_vm/loading.rs:13-226 */
int zkc_main_vm_initial_state(zkc_ctx *ctx, const zkc_vm_closed_form *io, const zkc_vm_isa *isa, zkc_vm_state *out);

/* Out-of-circuit run (the role of the external `zk_evm` crate + witness generation): executes `cycles` cycles of
 * `n_instances` independent VMs from bootloader start states (registers / flags may be preset), answering memory reads
 * from a per-instance memory model (code, stack, heap and aux heap pages of the root frame, 2^16 words each, code
 * pre-loaded from `code`, [n_instances][code_words][8] limbs; storage slots start at zero), and records what the
 * circuit needs: snapshots [n_instances][cycles + 1], witness [n_instances][cycles] and the popped frames
 * callstack_witness_out [n_instances][callstack_capacity] (n_callstack_out[n_instances] = how many were used).
 * The rollback queue is hash-chained BACKWARDS (every revertable log prepends to its frame's segment, log.rs:351-371,
 * and a reverting frame's segment must start where the forward queue ends, ret.rs:373-383), so the run takes two
 * passes: the first resolves every claimed rollback head / frame tail, the second records.  The resolved
 * rollback_queue_tail_for_block of each instance is written to rollback_tails_out [n_instances][4] (equal to the
 * start state's own tail when the root frame logs nothing) and snapshots[0] carries it.
 * Device buffers only (rollback_tails_out and n_callstack_out are host memory). */
int zkc_main_vm_simulate(zkc_ctx *ctx, const zkc_vm_isa *isa, const zkc_vm_state *initial_states, const uint32_t *code,
                         size_t code_words, size_t n_instances, size_t cycles, zkc_vm_state *snapshots_out,
                         zkc_vm_cycle_witness *witness_out, zkc_vm_callstack_witness *callstack_witness_out,
                         size_t callstack_capacity, uint32_t *n_callstack_out, uint64_t *rollback_tails_out,
                         zkc_status *status);

/* ---- generic allocation-check evaluator -----------------------------------------------------------------------------------
 * Every variable the reference allocates carries the range relation of its gadget type (Boolean::allocate -> x (x - 1) = 0,
 * UInt8 / UInt16 / UInt32::allocate_checked -> range-check lookups, Num -> a canonical field element).  This evaluator streams
 * ANY finished column-major trace [n_cols][rows] once (row pairs, 128-bit loads) and re-evaluates those relations from a
 * per-column class table; the circuits' row-local arithmetic relations have their own evaluators (zkc_ram_permutation_check_trace,
 * zkc_log_sorter_check_trace, zkc_main_vm_check_trace).  Returns the number of violating rows; status->first_bad_row and
 * status->failed_checks (bit k: a cell of class k is out of range) describe the first one; first_bad_column (may be NULL). */
enum zkc_col_class { ZKC_COL_FIELD = 0, ZKC_COL_BOOLEAN, ZKC_COL_U8, ZKC_COL_U16, ZKC_COL_U32, ZKC_COL_NUM_CLASSES };
int zkc_check_trace_columns(zkc_ctx *ctx, const uint64_t *trace, size_t n_cols, size_t rows, const uint8_t *col_class, int on_device,
                            uint64_t *violations, uint32_t *first_bad_column, zkc_status *status);

/* ---- linear_hasher: Keccak-256 of the L2 -> L1 message queue ---------------------------------------------
 * linear_hasher_entry_point, src/linear_hasher/mod.rs:35-214.  Every cycle pops one LogQuery (:107), serialises it into
 * L2_TO_L1_MESSAGE_BYTE_LENGTH = 88 bytes (ByteSerializable::into_bytes, base_structures/log_query/mod.rs:645-686:
 * shard_id, is_service, tx_number_in_block as 2 big-endian bytes, address 20, key 32, written_value 32, all big-endian),
 * appends them to a byte buffer (:116), absorbs 136 bytes + keccak-f[1600] when the buffer holds that many (:120-137) and
 * the padded remainder on the queue's last item (:142-168).  One instance per block: start_flag is enforced (:66).
 * The buffer length before cycle c is (88 c) mod 136: a compile-time constant of the cycle in the reference, a function of
 * the row index here. */
#define ZKC_LH_MESSAGE_BYTES 88
#define ZKC_KECCAK_RATE_BYTES 136
typedef struct zkc_linear_hasher_closed_form {
    uint32_t start_flag;      /* must be 1 (:66) */
    uint32_t completion_flag; /* out */
    zkc_queue_state4 queue_state;  /* LinearHasherInputData, input.rs:24-26 */
    uint32_t keccak256_hash[32];   /* LinearHasherOutputData (input.rs:39-41), one byte per element (out; expected if compare_expected) */
} zkc_linear_hasher_closed_form;

/* trace columns of one loop iteration (mod.rs:103-171) */
enum zkc_lh_col {
    ZKC_LH_QUEUE_IS_EMPTY = 0,  /* :104 */
    ZKC_LH_SHOULD_POP = 1,      /* :105 */
    ZKC_LH_ITEM = 2,            /* 36: popped record, flatten order log_query/mod.rs:62-101 */
    ZKC_LH_ENC = 38,            /* 20: LogQuery::encode (the pop's absorbed elements) */
    ZKC_LH_HEAD = 58,           /* 4: queue head after the pop */
    ZKC_LH_LEN = 62,            /* queue length after the pop */
    ZKC_LH_NOW_EMPTY = 63,      /* :109 */
    ZKC_LH_IS_LAST_SERIALIZATION = 64, /* :110 */
    ZKC_LH_BYTES = 65,          /* 88: into_bytes (:112) */
    ZKC_LH_CONTINUE_TO_ABSORB = 153, /* :118 */
    ZKC_LH_ABSORB_FULL = 154,   /* condition of the full-block absorption of this cycle (0 when the buffer stays short, :120) */
    ZKC_LH_ABSORB_LAST = 155,   /* absorb_as_last_round, :144-145 */
    ZKC_LH_STATE_MID = 156,     /* 50: keccak state after the conditional full-block round (:131-136): lane x + 5y as (low, high) u32 halves */
    ZKC_LH_STATE_OUT = 206,     /* 50: after the conditional last round (:162-167) */
    ZKC_LH_DONE = 256,          /* :170 */
    ZKC_LH_NUM_COLS = 257
};

#define ZKC_LH_CHK_START_FLAG (1u << 0)        /* :66 */
#define ZKC_LH_CHK_TRIVIAL_HEAD (1u << 1)      /* :71 */
#define ZKC_LH_CHK_TX_NUMBER_RANGE (1u << 2)   /* into_bytes: the two high bytes of tx_number_in_block are zero, log_query/mod.rs:666-668 */
#define ZKC_LH_CHK_QUEUE_CONSISTENCY (1u << 3) /* :173 */
#define ZKC_LH_CHK_NOT_COMPLETED (1u << 4)     /* :176: the queue must be empty after `limit` cycles */
#define ZKC_LH_CHK_QUEUE_HINT (1u << 5)        /* prev_tails is not the hash chain of the records */
#define ZKC_LH_CHK_STATE_HINT (1u << 6)        /* keccak_states is not the absorption chain */

/*   records, prev_tails : the queue's witness in pop order (CircuitQueueRawWitness, input.rs:78-87), as for the demultiplexer
 *   keccak_states       : NULL, or [limit][25] lanes: the keccak state AFTER every cycle (what an out-of-circuit run of the
 *                         hasher holds; the state before cycle 0 is zero).  Verified row by row -- every cycle is then
 *                         independent; when NULL the chain (one keccak-f per 136 bytes) is rebuilt sequentially on the device
 *   trace               : column-major [ZKC_LH_NUM_COLS][limit] or NULL */
int zkc_linear_hasher_entry_point(zkc_ctx *ctx, zkc_linear_hasher_closed_form *io, const zkc_log_query *records,
                                  const uint64_t *prev_tails, size_t n_records, const uint64_t *keccak_states, size_t limit,
                                  const zkc_sorter_options *options, int on_device, uint64_t *trace,
                                  uint64_t commitment[ZKC_COMMITMENT_LEN], zkc_status *status);

/* constraint evaluation of a finished linear_hasher trace: every relation of the loop (mod.rs:103-171) on every cycle, the keccak
 * sponge included (buffer = the BYTES columns of the two previous cycles).  gates: ZKC_GATES_GENERAL = everything but the Poseidon2
 * permutations of the pop; ZKC_GATES_ROUND_FUNCTION adds them; 0 = all. */
#define ZKC_LHV_BOOLEAN (1u << 0)        /* booleans, u32 / u8 ranges, field range of hash outputs */
#define ZKC_LHV_QUEUE (1u << 1)          /* is_empty / length / head bookkeeping of the popped queue */
#define ZKC_LHV_ENCODING (1u << 2)       /* LogQuery::encode, into_bytes */
#define ZKC_LHV_ROUND_FUNCTION (1u << 3)
#define ZKC_LHV_FLAGS (1u << 4)          /* now_empty, is_last_serialization, continue_to_absorb, the absorption conditions, done */
#define ZKC_LHV_ENFORCE (1u << 5)        /* tx_number_in_block fits two bytes */
#define ZKC_LHV_SPONGE (1u << 6)         /* keccak state after the conditional full-block and padded last rounds */
int zkc_linear_hasher_check_trace(zkc_ctx *ctx, const zkc_linear_hasher_closed_form *io, const uint64_t *trace, size_t limit, uint32_t gates,
                                  int on_device, uint64_t *violations, zkc_status *status);

#ifdef __cplusplus
}
#endif
#endif
