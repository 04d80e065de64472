"""ctypes view of include/zkc_b200.h (the C ABI of libzkc_b200.so) plus numpy record dtypes.

Nothing here computes: it only marshals the reference's witness structures
(/root/reference/src/ram_permutation/input.rs:27-116, src/fsm_input_output/mod.rs:32-48) into the
`#[repr(C)]` records the sm_100a engine consumes.  Loading fails loudly when the CUDA library is
missing -- there is no CPU fallback.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ZKC_B200_LIB") or os.path.join(_HERE, "libzkc_b200.so")  # the env override selects another BUILD of the same engine

GL_P = 0xFFFFFFFF00000001

ZKC_OK, ZKC_ERR_INVALID_ARGUMENT, ZKC_ERR_CUDA, ZKC_ERR_NO_DEVICE = 0, 1, 2, 3
ZKC_ERR_UNSATISFIED, ZKC_ERR_FSM_OUTPUT_MISMATCH, ZKC_ERR_QUEUE_WITNESS_INCONSISTENT = 4, 5, 6
CODE_NAMES = {0: "OK", 1: "INVALID_ARGUMENT", 2: "CUDA", 3: "NO_DEVICE", 4: "UNSATISFIED",
              5: "FSM_OUTPUT_MISMATCH", 6: "QUEUE_WITNESS_INCONSISTENT"}

MEMORY_QUERY_DTYPE = np.dtype([
    ("timestamp", "<u4"), ("memory_page", "<u4"), ("index", "<u4"), ("rw_flag", "<u4"), ("is_ptr", "<u4"),
    ("value", "<u4", (8,)), ("_pad", "<u4", (3,)),
])
assert MEMORY_QUERY_DTYPE.itemsize == 64


class Status(C.Structure):
    _fields_ = [("code", C.c_int32), ("cuda_error", C.c_int32), ("first_bad_row", C.c_int64),
                ("failed_checks", C.c_uint32), ("reserved", C.c_uint32)]


class QueueState12(C.Structure):
    _fields_ = [("head", C.c_uint64 * 12), ("tail", C.c_uint64 * 12), ("length", C.c_uint32), ("_pad", C.c_uint32)]


class QueueState4(C.Structure):
    _fields_ = [("head", C.c_uint64 * 4), ("tail", C.c_uint64 * 4), ("length", C.c_uint32), ("_pad", C.c_uint32)]


LOG_QUERY_DTYPE = np.dtype([
    ("address", "<u4", (5,)), ("key", "<u4", (8,)), ("read_value", "<u4", (8,)), ("written_value", "<u4", (8,)),
    ("tx_number_in_block", "<u4"), ("timestamp", "<u4"), ("flags", "<u4"),
])
assert LOG_QUERY_DTYPE.itemsize == 128


def lq_flags(aux=0, shard=0, rw=0, rollback=0, service=0):
    return int(aux) | int(shard) << 8 | int(rw) << 16 | int(rollback) << 17 | int(service) << 18


class LogQuery(C.Structure):
    _fields_ = [("address", C.c_uint32 * 5), ("key", C.c_uint32 * 8), ("read_value", C.c_uint32 * 8),
                ("written_value", C.c_uint32 * 8), ("tx_number_in_block", C.c_uint32), ("timestamp", C.c_uint32),
                ("flags", C.c_uint32)]


class EventsFsm(C.Structure):
    _fields_ = [("lhs_accumulator", C.c_uint64 * 2), ("rhs_accumulator", C.c_uint64 * 2),
                ("initial_unsorted_queue_state", QueueState4), ("intermediate_sorted_queue_state", QueueState4),
                ("final_result_queue_state", QueueState4), ("previous_key", C.c_uint32), ("_pad", C.c_uint32),
                ("previous_item", LogQuery)]


class EventsClosedForm(C.Structure):
    _fields_ = [("start_flag", C.c_uint32), ("completion_flag", C.c_uint32), ("initial_log_queue_state", QueueState4),
                ("intermediate_sorted_queue_state", QueueState4), ("final_queue_state", QueueState4),
                ("hidden_fsm_input", EventsFsm), ("hidden_fsm_output", EventsFsm)]


class StorageFsm(C.Structure):
    _fields_ = [("lhs_accumulator", C.c_uint64 * 2), ("rhs_accumulator", C.c_uint64 * 2),
                ("current_unsorted_queue_state", QueueState4), ("current_intermediate_sorted_queue_state", QueueState4),
                ("current_final_sorted_queue_state", QueueState4), ("cycle_idx", C.c_uint32),
                ("previous_packed_key", C.c_uint32 * 13), ("previous_key", C.c_uint32 * 8),
                ("previous_address", C.c_uint32 * 5), ("previous_timestamp", C.c_uint32),
                ("this_cell_has_explicit_read_and_rollback_depth_zero", C.c_uint32),
                ("this_cell_base_value", C.c_uint32 * 8), ("this_cell_current_value", C.c_uint32 * 8),
                ("this_cell_current_depth", C.c_uint32), ("_pad", C.c_uint32)]


class StorageClosedForm(C.Structure):
    _fields_ = [("start_flag", C.c_uint32), ("completion_flag", C.c_uint32), ("shard_id_to_process", C.c_uint32),
                ("_pad", C.c_uint32), ("unsorted_log_queue_state", QueueState4),
                ("intermediate_sorted_queue_state", QueueState4), ("final_sorted_queue_state", QueueState4),
                ("hidden_fsm_input", StorageFsm), ("hidden_fsm_output", StorageFsm)]


ST_COLS = dict(
    ORIGINAL_IS_EMPTY=0, SORTED_IS_EMPTY=1, SHOULD_POP=2, ORIGINAL_TIMESTAMP=3, UNSORTED_ITEM=4, UNSORTED_ENC=40,
    UNSORTED_EXT19=60, UNSORTED_HEAD=61, UNSORTED_LEN=65, SORTED_ITEM=66, SORTED_ENC=103, SORTED_HEAD=123, SORTED_LEN=127,
    SHARD_ID_IS_VALID=128, GP_CHAIN=129, GP_NEW=209, GP_ACC=213, CMP_DIFF=217, CMP_BORROW=230, CMP_LIMB_EQ=243,
    KEYS_ARE_EQUAL=256, PREVIOUS_KEY_IS_GREATER=257, TS_DIFF=258, PREVIOUS_TIMESTAMP_IS_LESS=259, MUST_ENFORCE=260,
    VALUE_IS_UNCHANGED=261, CURRENT_DEPTH_IS_ZERO=262, UNCHANGED_BUT_NOT_BY_ROLLBACK=263, ISSUE_PROTECTIVE_READ=264,
    SHOULD_WRITE=265, SHOULD_UPDATE=266, SHOULD_PUSH=267, NEW_NON_TRIVIAL_CELL=268, PUSH_ENC=269, PUSH_ROUND0=289,
    PUSH_ROUND1=301, PUSH_ROUND2=313, RESULT_TAIL=325, RESULT_LEN=329, CELL_BASE_VALUE=330, CELL_CURRENT_VALUE=338,
    CELL_CURRENT_DEPTH=346, CELL_HAS_READ_AT_DEPTH_ZERO=347, NON_TRIVIAL_AND_SAME_CELL=348, READ_OF_SAME_CELL=349,
    WRITE_OF_SAME_CELL=350, WRITE_NO_ROLLBACK=351, WRITE_ROLLBACK=352, READ_IS_EQUAL_TO_CURRENT=353,
    CHECK_READ_CONSISTENCY=354, ROLLBACK_DEPTH_IS_ZERO=355, READ_AT_DEPTH_ZERO_OF_SAME_CELL=356, NUM_COLS=357)
ST_CHK = dict(LENGTHS_EQUAL=1 << 0, EMPTY_SYNC=1 << 1, SHARD_ID=1 << 2, KEY_ORDER=1 << 3, TIMESTAMP_ORDER=1 << 4,
              FIRST_KEY_NONZERO=1 << 5, READ_CONSISTENCY=1 << 6, QUEUE_CONSISTENCY=1 << 7, GRAND_PRODUCT=1 << 8,
              TRIVIAL_HEAD=1 << 9, QUEUE_HINT=1 << 10, DEPTH_UNDERFLOW=1 << 11)


class KeccakFsm(C.Structure):
    _fields_ = [("read_precompile_call", C.c_uint32), ("read_unaligned_words_for_round", C.c_uint32),
                ("padding_round", C.c_uint32), ("completed", C.c_uint32), ("keccak_internal_state", C.c_uint8 * 200),
                ("timestamp_to_use_for_read", C.c_uint32), ("timestamp_to_use_for_write", C.c_uint32),
                ("input_page", C.c_uint32), ("input_memory_byte_offset", C.c_uint32), ("input_memory_byte_length", C.c_uint32),
                ("output_page", C.c_uint32), ("output_word_offset", C.c_uint32), ("needs_full_padding_round", C.c_uint32),
                ("buffer_bytes", C.c_uint8 * 192), ("buffer_filled", C.c_uint32), ("_pad", C.c_uint32),
                ("log_queue_state", QueueState4), ("memory_queue_state", QueueState12)]


class KeccakClosedForm(C.Structure):
    _fields_ = [("start_flag", C.c_uint32), ("completion_flag", C.c_uint32), ("initial_log_queue_state", QueueState4),
                ("initial_memory_queue_state", QueueState12), ("final_memory_state", QueueState12),
                ("hidden_fsm_input", KeccakFsm), ("hidden_fsm_output", KeccakFsm)]


class Sha256Fsm(C.Structure):
    _fields_ = [("read_precompile_call", C.c_uint32), ("read_words_for_round", C.c_uint32), ("completed", C.c_uint32),
                ("sha256_inner_state", C.c_uint32 * 8), ("timestamp_to_use_for_read", C.c_uint32),
                ("timestamp_to_use_for_write", C.c_uint32), ("input_page", C.c_uint32), ("input_offset", C.c_uint32),
                ("output_page", C.c_uint32), ("output_offset", C.c_uint32), ("num_rounds", C.c_uint32), ("_pad", C.c_uint32),
                ("log_queue_state", QueueState4), ("memory_queue_state", QueueState12)]


class Sha256ClosedForm(C.Structure):
    _fields_ = [("start_flag", C.c_uint32), ("completion_flag", C.c_uint32), ("initial_log_queue_state", QueueState4),
                ("initial_memory_queue_state", QueueState12), ("final_memory_state", QueueState12),
                ("hidden_fsm_input", Sha256Fsm), ("hidden_fsm_output", Sha256Fsm)]


SH_COLS = dict(FLAGS_IN=0, CALL_ITEM=3, REQ_HEAD=39, REQ_LEN=43, PARAMS=44, TS_READ=49, TS_WRITE=50, RESET_BUFFER=51,
               SHOULD_READ=52, QUERY=53, QUERY_STRIDE=22, MESSAGE=97, NUM_ROUNDS=113, STATE_IN=114, STATE_OUT=122,
               WRITE_RESULT=130, RESULT=131, WRITE_TAIL=139, WRITE_LEN=151, FLAGS_OUT=152, NUM_COLS=155)
SH_CHK_ZERO_ROUNDS = 1 << 7


class VmIsa(C.Structure):
    _fields_ = [("opcode_price", C.c_uint32 * 2048), ("opcode_props", C.c_uint64 * 2048), ("condition_table", (C.c_uint8 * 8) * 8),
                ("nop_opcode_encoding", C.c_uint64), ("panic_opcode_encoding", C.c_uint64), ("nop_bitspread", C.c_uint64),
                ("panic_bitspread", C.c_uint64), ("bootloader_base_page", C.c_uint32), ("bootloader_code_page", C.c_uint32),
                ("bootloader_calldata_page", C.c_uint32), ("starting_timestamp", C.c_uint32), ("starting_base_page", C.c_uint32),
                ("initial_frame_formal_eh_location", C.c_uint32), ("vm_initial_frame_ergs", C.c_uint32),
                ("bootloader_formal_address_low", C.c_uint32), ("bootloader_max_memory", C.c_uint32),
                ("vm_max_stack_depth", C.c_uint32), ("log_aux_bytes", C.c_uint32 * 4),
                ("initial_storage_write_pubdata_bytes", C.c_uint32), ("l1_message_pubdata_bytes", C.c_uint32),
                ("new_frame_memory_stipend", C.c_uint32), ("new_memory_pages_per_far_call", C.c_uint32),
                ("deployer_system_contract_address_low", C.c_uint32), ("ergs_per_code_word_decommittment", C.c_uint32),
                ("code_hash_version_byte", C.c_uint32), ("code_hash_yet_constructed_marker", C.c_uint32),
                ("code_hash_at_rest_marker", C.c_uint32), ("call_system_abi_registers", C.c_uint32 * 2),
                ("call_reserved_range", C.c_uint32 * 2), ("call_implicit_parameter_reg_idx", C.c_uint32)]


class VmRegister(C.Structure):
    _fields_ = [("is_pointer", C.c_uint32), ("value", C.c_uint32 * 8)]


class VmContext(C.Structure):
    _fields_ = [("this_address", C.c_uint32 * 5), ("caller", C.c_uint32 * 5), ("code_address", C.c_uint32 * 5),
                ("code_page", C.c_uint32), ("base_page", C.c_uint32), ("heap_upper_bound", C.c_uint32),
                ("aux_heap_upper_bound", C.c_uint32), ("reverted_queue_head", C.c_uint64 * 4),
                ("reverted_queue_tail", C.c_uint64 * 4), ("reverted_queue_segment_len", C.c_uint32), ("pc", C.c_uint32),
                ("sp", C.c_uint32), ("exception_handler_loc", C.c_uint32), ("ergs_remaining", C.c_uint32),
                ("is_static_execution", C.c_uint32), ("is_kernel_mode", C.c_uint32), ("this_shard_id", C.c_uint32),
                ("caller_shard_id", C.c_uint32), ("code_shard_id", C.c_uint32), ("context_u128_value_composite", C.c_uint32 * 4),
                ("is_local_call", C.c_uint32), ("log_queue_forward_part_length", C.c_uint32),
                ("log_queue_forward_tail", C.c_uint64 * 4)]


class VmState(C.Structure):
    _fields_ = [("previous_code_word", C.c_uint32 * 8), ("registers", VmRegister * 15), ("flags", C.c_uint32 * 3),
                ("timestamp", C.c_uint32), ("memory_page_counter", C.c_uint32), ("tx_number_in_block", C.c_uint32),
                ("previous_code_page", C.c_uint32), ("previous_super_pc", C.c_uint32), ("pending_exception", C.c_uint32),
                ("ergs_per_pubdata_byte", C.c_uint32), ("context_stack_depth", C.c_uint32), ("memory_queue_length", C.c_uint32),
                ("code_decommittment_queue_length", C.c_uint32), ("context_composite_u128", C.c_uint32 * 4), ("_pad", C.c_uint32),
                ("current_context", VmContext), ("stack_sponge_state", C.c_uint64 * 12), ("memory_queue_state", C.c_uint64 * 12),
                ("code_decommittment_queue_state", C.c_uint64 * 12)]


class VmCycleWitness(C.Structure):
    _fields_ = [("code_word", C.c_uint32 * 8), ("src0_is_pointer", C.c_uint32), ("src0_value", C.c_uint32 * 8),
                ("callstack_index", C.c_uint32), ("refund", C.c_uint32), ("suggested_page", C.c_uint32), ("value_a", C.c_uint32 * 8),
                ("value_b", C.c_uint32 * 8), ("rollback", C.c_uint64 * 4)]


class VmCallstackWitness(C.Structure):
    _fields_ = [("context", VmContext), ("previous_sponge_state", C.c_uint64 * 12)]


class VmClosedForm(C.Structure):
    _fields_ = [("start_flag", C.c_uint32), ("completion_flag", C.c_uint32), ("rollback_queue_tail_for_block", C.c_uint64 * 4),
                ("memory_queue_initial_tail", C.c_uint64 * 12), ("memory_queue_initial_length", C.c_uint32), ("_pad0", C.c_uint32),
                ("decommitment_queue_initial_tail", C.c_uint64 * 12), ("decommitment_queue_initial_length", C.c_uint32),
                ("zkporter_is_available", C.c_uint32), ("default_aa_code_hash", C.c_uint32 * 8),
                ("log_queue_final_state", QueueState4), ("memory_queue_final_state", QueueState12),
                ("decommitment_queue_final_state", QueueState12), ("hidden_fsm_input", VmState), ("hidden_fsm_output", VmState)]


class VmOptions(C.Structure):
    _fields_ = [("compare_expected", C.c_uint32), ("trace_layout", C.c_uint32), ("sponge_records_capacity", C.c_uint64),
                ("sponge_records", C.c_void_p)]


class VmColumns(C.Structure):
    """zkc_vm_columns: the per-cycle inputs as device-resident columns (struct of arrays)"""
    _fields_ = [("state_words", C.c_void_p), ("state_stride", C.c_size_t), ("witness_words", C.c_void_p), ("witness_stride", C.c_size_t)]


VM_STATE_WORDS, VM_WITNESS_WORDS = 294, 44


class VmSegmentHeader(C.Structure):
    """zkc_vm_segment_header: the first bytes of a segment blob of the input stream"""
    _fields_ = [(n, C.c_uint32) for n in (
        "magic", "first_cycle", "n_cycles", "n_dense_state", "n_dense_witness", "n_sparse_state", "n_sparse_witness", "reserved",
        "off_dense_state_word", "off_dense_witness_word", "off_dense_state", "off_dense_witness",
        "off_sparse_state_offsets", "off_sparse_state_index", "off_sparse_state_value",
        "off_sparse_witness_offsets", "off_sparse_witness_index", "off_sparse_witness_value", "blob_bytes", "reserved2")]


class VmInputSegment(C.Structure):
    _fields_ = [("blob", C.c_void_p), ("blob_bytes", C.c_uint64)]


class VmInputStream(C.Structure):
    _fields_ = [("limit", C.c_uint64), ("segment_cycles", C.c_uint32), ("n_segments", C.c_uint32), ("segments", C.POINTER(VmInputSegment))]


class VmPackedTrace(C.Structure):
    _fields_ = [("cols8", C.c_void_p), ("cols16", C.c_void_p), ("cols32", C.c_void_p), ("cols64", C.c_void_p),
                ("aux_records", C.c_void_p), ("aux_capacity", C.c_uint64), ("n_aux_records", C.c_uint64),
                ("sponge_records", C.c_void_p), ("sponge_capacity", C.c_uint64), ("n_sponge_records", C.c_uint64),
                ("limb_records", C.c_void_p), ("limb_capacity", C.c_uint64), ("n_limb_records", C.c_uint64)]


VM_SEGMENT_MAGIC = 0x5a4b5347
VM_TRACE_PACKED = 2
VM_PK_U8, VM_PK_U16, VM_PK_U32, VM_PK_U64, VM_PK_AUX_RECORD, VM_PK_SPONGE_RECORD, VM_PK_LIMB_RECORD = range(7)
VM_LIMB_CODE_WORD, VM_LIMB_SRC0_FROM_MEMORY, VM_LIMB_DST1 = range(3)
VM_LIMB_RECORD_DTYPE = np.dtype([("row", "<u4"), ("kind", "<u4"), ("v", "<u4", (9,)), ("reserved", "<u4")])
assert VM_LIMB_RECORD_DTYPE.itemsize == 48
VM_AUX_RECORD_DTYPE = np.dtype([("row", "<u4"), ("reserved", "<u4"), ("op_aux", "<u8", (48,)), ("queue_ends", "<u8", (10,))])
assert VM_AUX_RECORD_DTYPE.itemsize == 472 and C.sizeof(VmSegmentHeader) == 80
VM_SPONGE_RECORD_DTYPE = np.dtype([("row", "<u4"), ("slot", "<u4"), ("out", "<u8", (12,))])
assert VM_SPONGE_RECORD_DTYPE.itemsize == 104
VM_TRACE_DENSE, VM_TRACE_COMPACT = 0, 1


assert C.sizeof(VmState) == 1176 and C.sizeof(VmContext) == 240 and C.sizeof(VmIsa) == 24784
assert C.sizeof(VmClosedForm) == 3104 and C.sizeof(VmCycleWitness) == 176 and C.sizeof(VmCallstackWitness) == 336
VM_STATE_DTYPE = np.dtype((np.void, 1176))
VM_COLS = dict(
    SHOULD_SKIP_CYCLE=0, PENDING_EXCEPTION_IN=1, SHOULD_READ_OPCODE=2, SUPER_PC=3, SUB_PC=4, CODE_WORD=5, OPCODE=13, VARIANT=15,
    CONDITION_IDX=16, CONDITION=17, ERGS_COST=18, OUT_OF_ERGS=19, KERNEL_MODE_EXCEPTION=20, STATIC_EXCEPTION=21,
    CALLSTACK_IS_FULL=22, EXPLICIT_PANIC=23, MASK_INTO_PANIC=24, MASK_INTO_NOP=25, PROPS=26, DIRTY_ERGS_LEFT=27, SRC0_REG=28,
    SRC1_REG=29, DST0_REG=30, DST1_REG=31, IMM0=32, IMM1=33, SRC0_PAGE=34, SRC0_INDEX=35, SHOULD_READ_SRC0=36, SP_AFTER_SRC0=37,
    DST0_PAGE=38, DST0_INDEX=39, DST0_PERFORMS_MEMORY_ACCESS=40, NEW_SP=41, SRC0_FROM_MEMORY=42, SWAP_OPERANDS=51, SRC0=52,
    SRC1=61, DST0=70, DST1=79, PERFORM_DST0_MEMORY_WRITE=88, DST0_UPDATE_REGISTER=89, DST1_UPDATE_REGISTER=90, FLAGS_OUT=91,
    PENDING_EXCEPTION_OUT=94, PC_OUT=95, ERGS_OUT=96, HEAP_BOUND_OUT=97, AUX_HEAP_BOUND_OUT=98, MEMQ_LENGTH_OUT=99, DEPTH_OUT=100,
    FORWARD_TAIL_OUT=101, ROLLBACK_HEAD_OUT=106, SPONGE_ENFORCE=111, SPONGE_FINAL=120, OP_AUX=228, NUM_COLS=276)
VM_COMPACT_COLS, VM_COMPACT_OP_AUX = 276 - 117, 228 - 117
VMV = dict(BOOLEAN=1, RANGE=2, DECODE=4, EXCEPTION_MASKS=8, ADD_SUB=16, MUL_DIV=32, BINOP=64, FLAGS=128, SELECTION=256, SPONGE=512)


def vm_expand_compact_trace(compact, records, limit):
    """dense [.., 276, limit] trace from the COMPACT layout (159 columns + sponge records); compact: [159, limit] or
    [n, 159, limit]"""
    compact = np.asarray(compact)
    batched = compact.ndim == 3
    c3 = compact if batched else compact[None]
    n = c3.shape[0]
    dense = np.zeros((n, VM_COLS["NUM_COLS"], limit), dtype=np.uint64)
    dense[:, :VM_COLS["SPONGE_ENFORCE"]] = c3[:, :VM_COLS["SPONGE_ENFORCE"]]
    dense[:, VM_COLS["OP_AUX"]:] = c3[:, VM_COMPACT_OP_AUX:]
    rec = np.asarray(records)
    inst, row, slot = rec["row"] // limit, rec["row"] % limit, rec["slot"].astype(np.int64)
    dense[inst, VM_COLS["SPONGE_ENFORCE"] + slot, row] = 1
    for j in range(12):
        dense[inst, VM_COLS["SPONGE_FINAL"] + 12 * slot + j, row] = rec["out"][:, j]
    return dense if batched else dense[0]
VM_CHK = dict(INVALID_OPCODE=1, UNSUPPORTED_OPCODE=2, SNAPSHOT=4, DIV_RELATION=8, BOOTLOADER_EXIT=16, ROLLBACK_QUEUE=32,
              CALLSTACK=64, LOG_REFUND=128)
ZKC_ERR_UNSUPPORTED, ZKC_ERR_SNAPSHOT_MISMATCH = 7, 8
CODE_NAMES.update({7: "UNSUPPORTED", 8: "SNAPSHOT_MISMATCH"})


class PrecompileOptions(C.Structure):
    _fields_ = [("compare_expected", C.c_uint32), ("precompile_address", C.c_uint32), ("aux_byte", C.c_uint32),
                ("_pad", C.c_uint32)]


KC_COLS = dict(FLAGS_IN=0, CALL_ITEM=4, REQ_HEAD=40, REQ_LEN=44, PARAMS=45, TS_READ=51, TS_WRITE=52, RESET_BUFFER=53,
               READ_ZERO_LENGTH=54, READ_NON_ZERO_LENGTH=55, QUERY=56, QUERY_STRIDE=28, ZERO_BYTES_LEFT=224,
               CURRENTLY_FILLED=225, DO_ONE_BYTE_OF_PADDING=226, BUFFER_NOW_EMPTY=227, APPLY_PADDING=228, INPUT=229,
               STATE_OUT=365, WRITE_RESULT=565, RESULT=566, WRITE_TAIL=574, WRITE_LEN=586, FLAGS_OUT=587, BUFFER_OUT=591,
               NUM_COLS=783)
KC_CHK = dict(TRIVIAL_HEAD=1 << 0, AUX_BYTE=1 << 1, ADDRESS=1 << 2, BUFFER_OVERFLOW=1 << 3, QUEUE_CONSISTENCY=1 << 4,
              QUEUE_HINT=1 << 5, WITNESS_EXHAUSTED=1 << 6)
KECCAK256_PRECOMPILE_ADDRESS, SHA256_PRECOMPILE_ADDRESS, PRECOMPILE_AUX_BYTE = 0x8010, 0x02, 3


class SorterOptions(C.Structure):
    _fields_ = [("compare_expected", C.c_uint32), ("_pad", C.c_uint32 * 3)]


EV_COLS = dict(
    ORIGINAL_IS_EMPTY=0, SORTED_IS_EMPTY=1, SHOULD_POP=2, UNSORTED_ITEM=3, UNSORTED_ENC=39, UNSORTED_HEAD=59,
    UNSORTED_LEN=63, SORTED_ITEM=64, SORTED_ENC=100, SORTED_HEAD=120, SORTED_LEN=124, GP_CHAIN=125, GP_NEW=205,
    GP_ACC=209, CMP_DIFF=213, CMP_BORROW=214, KEYS_EQUAL=215, SAME_NONTRIVIAL_LOG=216, DIFFERENT_NONTRIVIAL_LOG=217,
    ITEM_KEYS_EQUAL=218, VALUES_EQUAL=219, SAME_BODY=220, PREVIOUS_IS_TRIVIAL=221, SHOULD_ENFORCE=222, MAYBE_ADD=223,
    ADD_TO_QUEUE=224, PUSH_ENC=225, PUSH_ROUND0=245, PUSH_ROUND1=257, PUSH_ROUND2=269, RESULT_TAIL=281, RESULT_LEN=285,
    NUM_COLS=286)
EV_CHK = dict(LENGTHS_EQUAL=1 << 0, EMPTY_SYNC=1 << 1, UNSORTED_IS_WRITE=1 << 2, SORTED_IS_WRITE=1 << 3, ORDER=1 << 4,
              NOT_ROLLBACK=1 << 5, IS_ROLLBACK=1 << 6, SAME_BODY=1 << 7, QUEUE_CONSISTENCY=1 << 8, GRAND_PRODUCT=1 << 9,
              TRIVIAL_HEAD=1 << 10, QUEUE_HINT=1 << 11)


DECOMMIT_QUERY_DTYPE = np.dtype([("code_hash", "<u4", (8,)), ("page", "<u4"), ("is_first", "<u4"), ("timestamp", "<u4"),
                                 ("_pad", "<u4")])
assert DECOMMIT_QUERY_DTYPE.itemsize == 48


class DecommitQuery(C.Structure):
    _fields_ = [("code_hash", C.c_uint32 * 8), ("page", C.c_uint32), ("is_first", C.c_uint32), ("timestamp", C.c_uint32),
                ("_pad", C.c_uint32)]


class DecommitSorterFsm(C.Structure):
    _fields_ = [("initial_queue_state", QueueState12), ("sorted_queue_state", QueueState12), ("final_queue_state", QueueState12),
                ("lhs_accumulator", C.c_uint64 * 2), ("rhs_accumulator", C.c_uint64 * 2),
                ("previous_packed_key", C.c_uint32 * 9), ("first_encountered_timestamp", C.c_uint32),
                ("previous_record", DecommitQuery)]


class DecommitSorterClosedForm(C.Structure):
    _fields_ = [("start_flag", C.c_uint32), ("completion_flag", C.c_uint32), ("initial_queue_state", QueueState12),
                ("sorted_queue_initial_state", QueueState12), ("final_queue_state", QueueState12),
                ("hidden_fsm_input", DecommitSorterFsm), ("hidden_fsm_output", DecommitSorterFsm)]


DQ_COLS = dict(
    ORIGINAL_IS_EMPTY=0, SORTED_IS_EMPTY=1, SHOULD_POP=2, UNSORTED_ITEM=3, UNSORTED_ENC=14, UNSORTED_HEAD=22, UNSORTED_LEN=34,
    SORTED_ITEM=35, SORTED_ENC=46, SORTED_HEAD=54, SORTED_LEN=66, GP_CHAIN=67, GP_NEW=99, GP_ACC=103, CMP_DIFF=107,
    CMP_BORROW=116, CMP_LIMB_EQ=125, KEYS_ARE_EQUAL=134, SAME_HASH=135, ENFORCE_MUST_BE_FIRST=136, PREVIOUS_IS_TRIVIAL=137,
    ENFORCE_SAME_MEMORY_PAGE=138, ADD_TO_QUEUE=139, PUSH_ITEM=140, PUSH_ENC=151, RESULT_TAIL=159, RESULT_LEN=171,
    FIRST_TIMESTAMP=172, NUM_COLS=173)
DQ_CHK = dict(LENGTHS_EQUAL=1 << 0, EMPTY_SYNC=1 << 1, ORDER=1 << 2, MUST_BE_FIRST=1 << 3, SAME_MEMORY_PAGE=1 << 4,
              QUEUE_CONSISTENCY=1 << 5, GRAND_PRODUCT=1 << 6, TRIVIAL_HEAD=1 << 7, QUEUE_HINT=1 << 8)


class DemuxFsm(C.Structure):
    _fields_ = [("initial_log_queue_state", QueueState4), ("output_queue_states", QueueState4 * 6)]


class DemuxClosedForm(C.Structure):
    _fields_ = [("start_flag", C.c_uint32), ("completion_flag", C.c_uint32), ("initial_log_queue_state", QueueState4),
                ("output_queue_states", QueueState4 * 6), ("hidden_fsm_input", DemuxFsm), ("hidden_fsm_output", DemuxFsm)]


class DemuxOptions(C.Structure):
    _fields_ = [("compare_expected", C.c_uint32), ("custom_constants", C.c_uint32), ("aux_bytes", C.c_uint32 * 4),
                ("precompile_addresses", C.c_uint32 * 3), ("_pad", C.c_uint32 * 3)]


DEMUX_QUEUES = ("storage", "events", "l1messages", "keccak256", "sha256", "ecrecover")  # enum LogType, demux_log_queue/mod.rs:224-232
STORAGE_AUX_BYTE, EVENT_AUX_BYTE, L1_MESSAGE_AUX_BYTE = 0, 1, 2
ECRECOVER_PRECOMPILE_ADDRESS = 0x01
DMX_COLS = dict(
    QUEUE_IS_EMPTY=0, EXECUTE=1, ITEM=2, ENC=38, HEAD=58, LEN=62, IS_AUX=63, IS_ADDRESS=67, IS_ROLLUP_SHARD=70,
    EXECUTE_PORTER_STORAGE=71, BITMASK=72, IS_BITMASK=78, EXEC_TAIL=79, EXEC_LEN=83, PUSH_ROUND0=84, PUSH_ROUND1=96,
    PUSH_ROUND2=108, QUEUE_TAILS=120, QUEUE_LENS=144, NUM_COLS=150)
DMX_CHK = dict(TRIVIAL_HEAD=1 << 0, PORTER_STORAGE=1 << 1, BITMASK=1 << 2, QUEUE_CONSISTENCY=1 << 3, QUEUE_HINT=1 << 4)


class LinearHasherClosedForm(C.Structure):
    """zkc_linear_hasher_closed_form: LinearHasherInputData / LinearHasherOutputData, linear_hasher/input.rs:24-41 (hidden FSM: `()`)"""
    _fields_ = [("start_flag", C.c_uint32), ("completion_flag", C.c_uint32), ("queue_state", QueueState4),
                ("keccak256_hash", C.c_uint32 * 32)]


LH_MESSAGE_BYTES, KECCAK_RATE_BYTES = 88, 136
LH_COLS = dict(QUEUE_IS_EMPTY=0, SHOULD_POP=1, ITEM=2, ENC=38, HEAD=58, LEN=62, NOW_EMPTY=63, IS_LAST_SERIALIZATION=64, BYTES=65,
               CONTINUE_TO_ABSORB=153, ABSORB_FULL=154, ABSORB_LAST=155, STATE_MID=156, STATE_OUT=206, DONE=256, NUM_COLS=257)
LH_CHK = dict(START_FLAG=1 << 0, TRIVIAL_HEAD=1 << 1, TX_NUMBER_RANGE=1 << 2, QUEUE_CONSISTENCY=1 << 3, NOT_COMPLETED=1 << 4,
              QUEUE_HINT=1 << 5, STATE_HINT=1 << 6)


class CodeDecommittmentFsm(C.Structure):
    _fields_ = [("sha256_inner_state", C.c_uint32 * 8), ("hash_to_compare_against", C.c_uint32 * 8), ("current_index", C.c_uint32),
                ("current_page", C.c_uint32), ("timestamp", C.c_uint32), ("num_rounds_left", C.c_uint32),
                ("length_in_bits", C.c_uint32), ("state_get_from_queue", C.c_uint32), ("state_decommit", C.c_uint32),
                ("finished", C.c_uint32)]


class CodeUnpackerFsm(C.Structure):
    _fields_ = [("internal_fsm", CodeDecommittmentFsm), ("decommittment_requests_queue_state", QueueState12),
                ("memory_queue_state", QueueState12)]


class CodeUnpackerClosedForm(C.Structure):
    _fields_ = [("start_flag", C.c_uint32), ("completion_flag", C.c_uint32), ("memory_queue_initial_state", QueueState12),
                ("sorted_requests_queue_initial_state", QueueState12), ("memory_queue_final_state", QueueState12),
                ("hidden_fsm_input", CodeUnpackerFsm), ("hidden_fsm_output", CodeUnpackerFsm)]


CODE_HASH_VERSION_TOP16 = 0x0100
CU_COLS = dict(
    FLAGS_IN=0, REQUEST=3, REQ_HEAD=14, REQ_LEN=26, VERSION_MATCHES=27, LENGTH_IN_WORDS=28, LENGTH_IN_ROUNDS=29, LENGTH_IN_BITS=30,
    TIMESTAMP=31, PAGE=32, HASH_TO_COMPARE=33, DECOMMIT=41, NUM_ROUNDS_LEFT=42, LAST_ROUND=43, FINALIZE=44, PROCESS_SECOND_WORD=45,
    WORD0=46, WORD1=54, INDEX0=62, INDEX1=63, INDEX_OUT=64, MEM_TAIL0=65, MEM_TAIL1=78, MESSAGE=91, STATE_IN=107, STATE_NEW=115,
    STATE_OUT=123, FLAGS_OUT=131, NUM_COLS=134)
CU_CHK = dict(VERSION=1 << 0, LENGTH=1 << 1, HASH=1 << 2, QUEUE_CONSISTENCY=1 << 3, QUEUE_HINT=1 << 4, WITNESS_EXHAUSTED=1 << 5)


class RamInputData(C.Structure):
    _fields_ = [("unsorted_queue_initial_state", QueueState12), ("sorted_queue_initial_state", QueueState12),
                ("non_deterministic_bootloader_memory_snapshot_length", C.c_uint32), ("_pad", C.c_uint32)]


class RamFsm(C.Structure):
    _fields_ = [("lhs_accumulator", C.c_uint64 * 2), ("rhs_accumulator", C.c_uint64 * 2),
                ("current_unsorted_queue_state", QueueState12), ("current_sorted_queue_state", QueueState12),
                ("previous_sorting_key", C.c_uint32 * 3), ("previous_full_key", C.c_uint32 * 2),
                ("previous_value", C.c_uint32 * 8), ("previous_is_ptr", C.c_uint32),
                ("num_nondeterministic_writes", C.c_uint32), ("_pad", C.c_uint32)]


class RamClosedForm(C.Structure):
    _fields_ = [("start_flag", C.c_uint32), ("completion_flag", C.c_uint32), ("observable_input", RamInputData),
                ("hidden_fsm_input", RamFsm), ("hidden_fsm_output", RamFsm)]


class RamOptions(C.Structure):
    _fields_ = [("bootloader_heap_page", C.c_uint32), ("compare_expected", C.c_uint32), ("_pad", C.c_uint32 * 2)]


# trace columns, enum zkc_ram_col
RAM_COLS = dict(
    UNSORTED_IS_EMPTY=0, SORTED_IS_EMPTY=1, CAN_POP=2, UNSORTED_ITEM=3, UNSORTED_ENC=16, UNSORTED_HEAD=24,
    UNSORTED_LEN=36, SORTED_ITEM=37, SORTED_ENC=50, SORTED_HEAD=58, SORTED_LEN=70, TS_IS_ZERO=71,
    PAGE_IS_BOOTLOADER_HEAP=72, IS_NONDET_WRITE=73, NUM_NONDET_WRITES=74, CMP_DIFF=75, CMP_BORROW=78,
    CMP_LIMB_EQ=81, KEYS_EQUAL=84, PREV_KEY_SMALLER=85, SAME_CELL=86, VALUE_EQUAL=87, VALUE_IS_ZERO=88,
    IS_ZERO=89, PTR_EQUALITY=90, VALUE_AND_PTR_EQUAL=91, READ_UNINIT=92, CHECK_EQUALITY=93, GP_CHAIN=94,
    GP_NEW=126, GP_ACC=130, UNSORTED_ENC_BYTES=134, SORTED_ENC_BYTES=146, UNSORTED_LEN_INV=158, SORTED_LEN_INV=159, TS_INV=160,
    PAGE_DIFF=161, PAGE_DIFF_INV=162, CMP_DIFF_INV=163, CELL_DIFF=166, CELL_DIFF_INV=168, CELL_LIMB_EQ=170, VALUE_DIFF=172,
    VALUE_DIFF_INV=180, VALUE_LIMB_EQ=188, VALUE_ZERO_DIFF=196, VALUE_ZERO_DIFF_INV=204, VALUE_ZERO_LIMB_EQ=212, PTR_DIFF=220,
    PTR_DIFF_INV=221, NUM_COLS=222)

RAM_CHK = dict(LENGTHS_EQUAL=1 << 0, EMPTY_SYNC=1 << 1, ASCENDING=1 << 2, UNINIT_READ_ZERO=1 << 3,
               READ_CONSISTENT=1 << 4, QUEUE_CONSISTENCY=1 << 5, GRAND_PRODUCT=1 << 6, NONDET_COUNT=1 << 7,
               TRIVIAL_HEAD=1 << 8, RANGE=1 << 9, QUEUE_HINT=1 << 10)

_u64p = C.POINTER(C.c_uint64)
_vp = C.c_void_p

# name -> (restype, argtypes); every symbol include/zkc_b200.h declares
SIGNATURES = {
    "zkc_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "zkc_destroy": (None, [_vp]),
    "zkc_set_stream": (C.c_int, [_vp, _vp]),
    "zkc_version": (C.c_char_p, []),
    "zkc_launch_count": (C.c_uint64, [_vp]),
    "zkc_sm_count": (C.c_int, [_vp]),
    "zkc_profile_enable": (C.c_int, [_vp, C.c_int]),
    "zkc_profile_query": (C.c_int, [_vp, C.c_char_p, C.POINTER(C.c_double), _u64p]),
    "zkc_profile_reset": (C.c_int, [_vp]),
    "zkc_host_alloc": (_vp, [C.c_size_t]),
    "zkc_host_free": (None, [_vp]),
    "zkc_poseidon2_permute": (C.c_int, [_vp, _vp, _vp, C.c_size_t, C.c_int]),
    "zkc_commit_encoding": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t, _vp, C.c_int]),
    "zkc_field_ops": (C.c_int, [_vp, _vp, _vp, _vp, C.c_size_t, _vp, _vp, _vp, _vp]),
    "zkc_accumulate_grand_products": (C.c_int, [_vp, _vp, _vp, _vp, C.c_size_t, C.c_size_t, _vp, _vp, _vp, _vp, _vp, C.c_int]),
    "zkc_check_trace_columns": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t, _vp, C.c_int, _u64p, C.POINTER(C.c_uint32), C.POINTER(Status)]),
    "zkc_scale_accumulators": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t, _vp, C.c_int]),
    "zkc_memory_queue_simulate": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t, _vp, _vp, C.c_int]),
    "zkc_log_queue_simulate": (C.c_int, [_vp, _vp, _vp, C.c_size_t, C.c_size_t, _vp, _vp, C.c_int]),
    "zkc_log_sorter_entry_point": (C.c_int, [_vp, C.POINTER(EventsClosedForm), _vp, _vp, C.c_size_t, _vp, _vp, C.c_size_t,
                                             _vp, C.c_size_t, C.c_size_t, C.POINTER(SorterOptions), C.c_int, _vp, _vp,
                                             C.POINTER(Status)]),
    "zkc_log_sorter_check_trace": (C.c_int, [_vp, C.POINTER(EventsClosedForm), _vp, C.c_size_t, C.c_uint32, C.c_int, _u64p, C.POINTER(Status)]),
    "zkc_storage_validity_check_trace": (C.c_int, [_vp, C.POINTER(StorageClosedForm), _vp, C.c_size_t, C.c_uint32, C.c_int, _u64p, C.POINTER(Status)]),
    "zkc_keccak256_round_function_check_trace": (C.c_int, [_vp, C.POINTER(KeccakClosedForm), C.POINTER(PrecompileOptions), _vp, C.c_size_t, C.c_uint32, C.c_int, _u64p, C.POINTER(Status)]),
    "zkc_linear_hasher_check_trace": (C.c_int, [_vp, C.POINTER(LinearHasherClosedForm), _vp, C.c_size_t, C.c_uint32, C.c_int, _u64p, C.POINTER(Status)]),
    "zkc_code_unpacker_check_trace": (C.c_int, [_vp, C.POINTER(CodeUnpackerClosedForm), _vp, C.c_size_t, C.c_uint32, C.c_int, _u64p, C.POINTER(Status)]),
    "zkc_sha256_round_function_check_trace": (C.c_int, [_vp, C.POINTER(Sha256ClosedForm), C.POINTER(PrecompileOptions), _vp, C.c_size_t, C.c_uint32, C.c_int, _u64p, C.POINTER(Status)]),
    "zkc_demux_log_queue_check_trace": (C.c_int, [_vp, C.POINTER(DemuxClosedForm), C.POINTER(DemuxOptions), _vp, C.c_size_t, C.c_uint32, C.c_int, _u64p, C.POINTER(Status)]),
    "zkc_sort_decommittments_check_trace": (C.c_int, [_vp, C.POINTER(DecommitSorterClosedForm), _vp, C.c_size_t, C.c_uint32, C.c_int, _u64p, C.POINTER(Status)]),
    "zkc_decommit_queue_simulate": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t, _vp, _vp, C.c_int]),
    "zkc_sort_decommittments_entry_point": (C.c_int, [_vp, C.POINTER(DecommitSorterClosedForm), _vp, _vp, C.c_size_t, _vp, _vp,
                                                      C.c_size_t, _vp, C.c_size_t, C.c_size_t, C.POINTER(SorterOptions),
                                                      C.c_int, _vp, _vp, C.POINTER(Status)]),
    "zkc_demux_log_queue_entry_point": (C.c_int, [_vp, C.POINTER(DemuxClosedForm), _vp, _vp, C.c_size_t, _vp,
                                                  C.POINTER(C.c_size_t), C.c_size_t, C.POINTER(DemuxOptions), C.c_int, _vp, _vp,
                                                  C.POINTER(Status)]),
    "zkc_linear_hasher_entry_point": (C.c_int, [_vp, C.POINTER(LinearHasherClosedForm), _vp, _vp, C.c_size_t, _vp, C.c_size_t,
                                                C.POINTER(SorterOptions), C.c_int, _vp, _vp, C.POINTER(Status)]),
    "zkc_code_unpacker_entry_point": (C.c_int, [_vp, C.POINTER(CodeUnpackerClosedForm), _vp, _vp, C.c_size_t, _vp, C.c_size_t, _vp,
                                                C.c_size_t, C.c_size_t, C.POINTER(SorterOptions), C.c_int, _vp, _vp, C.POINTER(Status)]),
    "zkc_storage_validity_entry_point": (C.c_int, [_vp, C.POINTER(StorageClosedForm), _vp, _vp, C.c_size_t, _vp, _vp, _vp,
                                                   C.c_size_t, _vp, C.c_size_t, C.c_size_t, C.POINTER(SorterOptions),
                                                   C.c_int, _vp, _vp, C.POINTER(Status)]),
    "zkc_keccak256_round_function_entry_point": (C.c_int, [_vp, C.POINTER(KeccakClosedForm), _vp, _vp, C.c_size_t, _vp,
                                                           C.c_size_t, _vp, C.c_size_t, C.c_size_t,
                                                           C.POINTER(PrecompileOptions), C.c_int, _vp, _vp,
                                                           C.POINTER(Status)]),
    "zkc_sha256_round_function_entry_point": (C.c_int, [_vp, C.POINTER(Sha256ClosedForm), _vp, _vp, C.c_size_t, _vp,
                                                        C.c_size_t, _vp, C.c_size_t, C.c_size_t,
                                                        C.POINTER(PrecompileOptions), C.c_int, _vp, _vp, C.POINTER(Status)]),
    "zkc_main_vm_entry_point": (C.c_int, [_vp, C.POINTER(VmClosedForm), C.POINTER(VmIsa), _vp, _vp, _vp, C.c_size_t, C.c_size_t,
                                          C.POINTER(VmOptions), C.c_int, _vp, _vp, C.POINTER(Status)]),
    "zkc_main_vm_entry_point_batch": (C.c_int, [_vp, _vp, C.c_size_t, C.POINTER(VmIsa), _vp, _vp, _vp, C.c_size_t, C.c_size_t,
                                                C.POINTER(VmOptions), C.c_int, _vp, _vp, _vp]),
    "zkc_main_vm_entry_point_columns": (C.c_int, [_vp, _vp, C.c_size_t, C.POINTER(VmIsa), C.POINTER(VmColumns), _vp, C.c_size_t, C.c_size_t,
                                                  C.POINTER(VmOptions), C.c_int, _vp, _vp, _vp]),
    "zkc_vm_encode_input_stream": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t, C.POINTER(C.POINTER(VmInputStream)), _u64p]),
    "zkc_vm_input_stream_free": (None, [C.POINTER(VmInputStream)]),
    "zkc_vm_packed_layout": (None, [_vp, _vp, _vp]),
    "zkc_main_vm_entry_point_stream": (C.c_int, [_vp, _vp, C.c_size_t, C.POINTER(VmIsa), _vp, _vp, C.c_size_t, C.c_size_t,
                                                 C.POINTER(VmOptions), C.POINTER(VmPackedTrace), _vp, _vp]),
    "zkc_main_vm_rows_to_columns": (C.c_int, [_vp, _vp, _vp, C.c_size_t, C.c_size_t, _vp, C.c_size_t, _vp, C.c_size_t]),
    "zkc_main_vm_gadget_cells": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t, C.c_int, _vp]),
    "zkc_main_vm_state_gadget_cells": (C.c_int, [_vp, _vp, _vp, C.c_size_t, C.c_size_t, C.c_int, _vp]),
    "zkc_main_vm_memory_sponge_cells": (C.c_int, [_vp, _vp, _vp, C.c_size_t, C.c_size_t, C.c_int, _vp]),
    "zkc_main_vm_prestate_cells": (C.c_int, [_vp, _vp, _vp, C.c_size_t, C.c_size_t, C.c_int, _vp]),
    "zkc_main_vm_writeback_cells": (C.c_int, [_vp, C.POINTER(VmIsa), _vp, _vp, C.c_size_t, C.c_size_t, C.c_int, _vp]),
    "zkc_main_vm_check_trace": (C.c_int, [_vp, C.POINTER(VmIsa), _vp, C.c_size_t, C.c_size_t, C.c_int, _u64p, C.POINTER(Status)]),
    "zkc_main_vm_initial_state": (C.c_int, [_vp, C.POINTER(VmClosedForm), C.POINTER(VmIsa), C.POINTER(VmState)]),
    "zkc_main_vm_simulate": (C.c_int, [_vp, C.POINTER(VmIsa), _vp, _vp, C.c_size_t, C.c_size_t, C.c_size_t, _vp, _vp, _vp,
                                       C.c_size_t, _vp, _vp, C.POINTER(Status)]),
    "zkc_ram_permutation_entry_point": (C.c_int, [_vp, C.POINTER(RamClosedForm), _vp, _vp, C.c_size_t, _vp, _vp,
                                                  C.c_size_t, C.c_size_t, C.POINTER(RamOptions), C.c_int, _vp, _vp,
                                                  C.POINTER(Status)]),
    "zkc_ram_permutation_check_trace": (C.c_int, [_vp, C.POINTER(RamClosedForm), _vp, C.c_size_t, C.POINTER(RamOptions),
                                                  C.c_uint32, C.c_int, _u64p, C.POINTER(Status)]),
}

GATES_GENERAL, GATES_ROUND_FUNCTION = 1, 2


def _gadget_columns(macro="ZKC_VM_GADGET_COLUMNS", enum="zkc_vm_gadget_col"):
    """An X-macro column list of include/zkc_b200.h (the list is the single source): name -> first column, + NUM_COLS"""
    import re
    text = open(os.path.join(os.path.dirname(_HERE), "include", "zkc_b200.h")).read()
    body = text[text.index(f"#define {macro}(X)"):text.index(f"enum {enum} {{")]
    cols, widths, n = {}, {}, 0
    for name, width in re.findall(r"X\((\w+), (\d+)\)", body):
        cols[name], widths[name] = n, int(width)
        n += int(width)
    cols["NUM_COLS"] = n
    return cols, widths


VMG_COLS, VMG_WIDTHS = _gadget_columns()
VMS_COLS, VMS_WIDTHS = _gadget_columns("ZKC_VM_STATE_GADGET_COLUMNS", "zkc_vm_state_gadget_col")  # ptr / jump / context block
VMQ_COLS, VMQ_WIDTHS = _gadget_columns("ZKC_VM_MEMORY_SPONGE_COLUMNS", "zkc_vm_memory_sponge_col")  # fetch / src0 / dst0 memory-queue relations
VMW_COLS, VMW_WIDTHS = _gadget_columns("ZKC_VM_WRITEBACK_COLUMNS", "zkc_vm_writeback_col")  # register write-back of the state diffs
VMP_COLS, VMP_WIDTHS = _gadget_columns("ZKC_VM_PRESTATE_COLUMNS", "zkc_vm_prestate_col")  # create_prestate: selectors, select chains, locations, swap
EVV = dict(BOOLEAN=1 << 0, QUEUE_LEN=1 << 1, ENCODING=1 << 2, ROUND_FUNCTION=1 << 3, COMPARISON=1 << 4, FLAGS=1 << 5, ENFORCE=1 << 6,
           GP_CHAIN=1 << 7, GP_ACC=1 << 8, RESULT_QUEUE=1 << 9)
KCV = dict(BOOLEAN=1 << 0, QUEUE=1 << 1, FSM=1 << 2, ROUND_FUNCTION=1 << 3, PARAMS=1 << 4, ENFORCE=1 << 5, SPONGE=1 << 6, MEMORY_QUEUE=1 << 7,
           QUERIES=1 << 8, BUFFER=1 << 9)  # ZKC_KCV_*
LHV = dict(BOOLEAN=1 << 0, QUEUE=1 << 1, ENCODING=1 << 2, ROUND_FUNCTION=1 << 3, FLAGS=1 << 4, ENFORCE=1 << 5, SPONGE=1 << 6)  # ZKC_LHV_*
CUV = dict(BOOLEAN=1 << 0, QUEUE=1 << 1, FSM=1 << 2, ROUND_FUNCTION=1 << 3, LENGTH=1 << 4, ENFORCE=1 << 5, COMPRESSION=1 << 6, MEMORY_QUEUE=1 << 7,
           SELECTS=1 << 8)  # ZKC_CUV_*
SHV = dict(BOOLEAN=1 << 0, QUEUE=1 << 1, FSM=1 << 2, ROUND_FUNCTION=1 << 3, PARAMS=1 << 4, ENFORCE=1 << 5, COMPRESSION=1 << 6, MEMORY_QUEUE=1 << 7)  # ZKC_SHV_*
DMXV = dict(BOOLEAN=1 << 0, QUEUE_LEN=1 << 1, ENCODING=1 << 2, ROUND_FUNCTION=1 << 3, FLAGS=1 << 4, ENFORCE=1 << 5, OUTPUT_QUEUES=1 << 6)  # ZKC_DMXV_*
DQV = dict(BOOLEAN=1 << 0, QUEUE_LEN=1 << 1, ENCODING=1 << 2, ROUND_FUNCTION=1 << 3, COMPARISON=1 << 4, FLAGS=1 << 5, ENFORCE=1 << 6,
           GP_CHAIN=1 << 7, GP_ACC=1 << 8, RESULT_QUEUE=1 << 9)  # ZKC_DQV_*
STV = dict(BOOLEAN=1 << 0, QUEUE_LEN=1 << 1, ENCODING=1 << 2, ROUND_FUNCTION=1 << 3, COMPARISON=1 << 4, FLAGS=1 << 5, ENFORCE=1 << 6,
           GP_CHAIN=1 << 7, GP_ACC=1 << 8, RESULT_QUEUE=1 << 9, CELL_STATE=1 << 10)  # ZKC_STV_*
RAMV = dict(BOOLEAN=1 << 0, QUEUE_LEN=1 << 1, ENCODING=1 << 2, ROUND_FUNCTION=1 << 3, NONDET=1 << 4, COMPARISON=1 << 5,
            FLAGS=1 << 6, ENFORCE=1 << 7, GP_CHAIN=1 << 8, GP_ACC=1 << 9, GADGET_CELLS=1 << 10)


def load_library(path=LIB_PATH):
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: build the sm_100a engine first (python -c 'import __graft_entry__ as g; g.build()'). "
            "There is no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library drift apart
        fn.restype = res
        fn.argtypes = args
    return lib
