"""B200-native witness-generation / constraint-evaluation engine for the data-parallel hot path of
the zkSync Era zkEVM circuits (reference: matter-labs/era-zkevm_circuits).  The compute lives in
csrc/ (hand-written sm_100a CUDA behind the C ABI of include/zkc_b200.h); this package is the
host-side mirror of the reference's entry points."""
from . import abi  # noqa: F401
from . import wire  # noqa: F401
from .engine import Engine, ZkcError  # noqa: F401
from .ram_permutation import (  # noqa: F401
    RamPermutationCircuitInstanceWitness,
    RamPermutationResult,
    ram_permutation_check_trace,
    ram_permutation_entry_point,
)
from .log_sorter import (  # noqa: F401
    EventsDeduplicatorInstanceWitness,
    SorterResult,
    log_sorter_check_trace,
    sort_and_deduplicate_events_entry_point,
)
from .storage_validity import (  # noqa: F401
    StorageDeduplicatorInstanceWitness,
    sort_and_deduplicate_storage_access_entry_point,
    storage_validity_check_trace,
)
from .sort_decommittment_requests import (  # noqa: F401
    CodeDecommittmentsDeduplicatorInstanceWitness,
    sort_and_deduplicate_code_decommittments_entry_point,
    sort_decommittments_check_trace,
)
from .demux_log_queue import (  # noqa: F401
    LogDemuxerCircuitInstanceWitness,
    demultiplex_storage_logs_enty_point,
    demux_log_queue_check_trace,
)
from .linear_hasher import (  # noqa: F401
    LinearHasherCircuitInstanceWitness,
    linear_hasher_entry_point,
    linear_hasher_check_trace,
)
from .code_unpacker_sha256 import (  # noqa: F401
    CodeDecommitterCircuitInstanceWitness,
    unpack_code_into_memory_entry_point,
    code_unpacker_check_trace,
)
from .keccak256_round_function import (  # noqa: F401
    Keccak256RoundFunctionCircuitInstanceWitness,
    keccak256_round_function_entry_point,
    keccak256_round_function_check_trace,
)
from .sha256_round_function import (  # noqa: F401
    Sha256RoundFunctionCircuitInstanceWitness,
    sha256_round_function_entry_point,
    sha256_round_function_check_trace,
)
from .main_vm import (  # noqa: F401
    VmCircuitWitness,
    main_vm_check_trace,
    main_vm_gadget_cells,
    main_vm_state_gadget_cells,
    main_vm_memory_sponge_cells,
    main_vm_prestate_cells,
    main_vm_writeback_cells,
    main_vm_entry_point,
    main_vm_entry_point_batch,
    main_vm_entry_point_columns,
    main_vm_rows_to_columns,
    VmColumnInputs,
    VmPackedTraceBuffers,
    main_vm_entry_point_stream,
    vm_decode_input_stream,
    vm_encode_input_stream,
    vm_expand_packed_trace,
    vm_packed_layout,
    vm_packed_trace_buffers,
    main_vm_initial_state,
    main_vm_simulate,
)
