// zkc_b200.hpp -- header-only C++ host side above the C ABI (include/zkc_b200.h).
//
// The reference (matter-labs/era-zkevm_circuits) is compiled code (Rust) whose toolchain is not part of this image, so the
// host mirror of its entry points that a native caller links is C++: one function per reference `*_entry_point`, with the
// SAME NAME and argument meaning (`fn(cs, witness, round_function, limit) -> [Num<F>; 4]`: the constraint system and the
// round function are the engine, the witness struct and `limit` are the arguments, the commitment comes back in the
// result together with everything else the reference computes), and the same failure behaviour: where the reference
// panics (hook_compare_witness, an exhausted witness deque, an inconsistent queue witness) these functions THROW
// (zkc_b200::Error); an unsatisfiable constraint system is reported in Result::status unless `throw_if_unsatisfied`.
// Witness structs mirror the reference's `*CircuitInstanceWitness` (deques of (item, previous queue state) pairs kept as
// two parallel vectors, which is how the C ABI takes them).  Host buffers only; callers that keep data in HBM use the
// C ABI directly.  There is no CPU fallback: without a device Engine's constructor throws (ZKC_ERR_NO_DEVICE).
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "zkc_b200.h"

namespace zkc_b200 {

using Commitment = std::array<uint64_t, ZKC_COMMITMENT_LEN>;
using State12 = std::array<uint64_t, ZKC_FULL_STATE>;
using State4 = std::array<uint64_t, ZKC_QUEUE_STATE>;

inline const char *code_name(int code) {
    switch (code) {
        case ZKC_OK: return "OK";
        case ZKC_ERR_INVALID_ARGUMENT: return "INVALID_ARGUMENT";
        case ZKC_ERR_CUDA: return "CUDA";
        case ZKC_ERR_NO_DEVICE: return "NO_DEVICE";
        case ZKC_ERR_UNSATISFIED: return "UNSATISFIED";
        case ZKC_ERR_FSM_OUTPUT_MISMATCH: return "FSM_OUTPUT_MISMATCH";
        case ZKC_ERR_QUEUE_WITNESS_INCONSISTENT: return "QUEUE_WITNESS_INCONSISTENT";
        default: return "UNKNOWN";
    }
}

struct Error : std::runtime_error {
    int code;
    zkc_status status;
    Error(const std::string &what, int code_, const zkc_status &st)
        : std::runtime_error(what + ": " + code_name(code_) + " (first_bad_row=" + std::to_string((long long)st.first_bad_row) +
                             ", failed_checks=" + std::to_string(st.failed_checks) + ", cuda=" + std::to_string(st.cuda_error) + ")"),
          code(code_), status(st) {}
};

// one engine per GPU / process (zkc_create)
class Engine {
  public:
    explicit Engine(int device = 0) {
        const int rc = zkc_create(device, &h_);
        if (rc != ZKC_OK) throw Error("zkc_create", rc, zkc_status{rc, 0, -1, 0, 0});
    }
    ~Engine() { if (h_) zkc_destroy(h_); }
    Engine(const Engine &) = delete;
    Engine &operator=(const Engine &) = delete;
    zkc_ctx *handle() const { return h_; }
    uint64_t launch_count() const { return zkc_launch_count(h_); }
    static std::string version() { return zkc_version(); }

  private:
    zkc_ctx *h_ = nullptr;
};

// what a reference entry point computes: the commitment (its return value), the closed-form input with outputs filled
// in, the witness values of every loop iteration (column-major, `columns` x `limit`), and which enforcement failed
template <class ClosedForm>
struct Result {
    Commitment commitment{};
    ClosedForm closed_form_input{};
    std::vector<uint64_t> trace;
    size_t columns = 0, limit = 0;
    zkc_status status{ZKC_OK, 0, -1, 0, 0};
    uint64_t cell(size_t column, size_t row) const { return trace[column * limit + row]; }
};

namespace detail {
template <class R>
inline void finish(const char *what, int rc, const R &r, bool throw_if_unsatisfied) {
    const bool panic = rc == ZKC_ERR_INVALID_ARGUMENT || rc == ZKC_ERR_CUDA || rc == ZKC_ERR_NO_DEVICE ||
                       rc == ZKC_ERR_FSM_OUTPUT_MISMATCH || rc == ZKC_ERR_QUEUE_WITNESS_INCONSISTENT;
    if (panic || (rc != ZKC_OK && throw_if_unsatisfied)) throw Error(what, rc, r.status);
}
template <class T, size_t N>
inline const uint64_t *flat(const std::vector<std::array<T, N>> &v) { return v.empty() ? nullptr : v.front().data(); }
template <class T>
inline const T *ptr(const std::vector<T> &v) { return v.empty() ? nullptr : v.data(); }
template <class R>
inline uint64_t *make_trace(R &r, size_t columns, size_t limit, bool want) {
    r.columns = columns; r.limit = limit;
    if (!want) return nullptr;
    r.trace.assign(columns * limit + 1, 0);
    return r.trace.data();
}
inline void same_length(const char *what, size_t a, size_t b) {
    if (a != b) throw std::invalid_argument(std::string(what) + ": a queue witness and its previous-state column differ in length");
}
}  // namespace detail

// ---- ram_permutation -----------------------------------------------------------------------------------------------
// RamPermutationCircuitInstanceWitness, /root/reference/src/ram_permutation/input.rs:99-116
struct RamPermutationCircuitInstanceWitness {
    zkc_ram_closed_form closed_form_input{};
    std::vector<zkc_memory_query> unsorted_queue_witness;
    std::vector<State12> unsorted_queue_prev_states;
    std::vector<zkc_memory_query> sorted_queue_witness;
    std::vector<State12> sorted_queue_prev_states;
};

// ram_permutation_entry_point, /root/reference/src/ram_permutation/mod.rs:31-210
inline Result<zkc_ram_closed_form> ram_permutation_entry_point(Engine &e, const RamPermutationCircuitInstanceWitness &w, size_t limit,
                                                               bool want_trace = true, const zkc_ram_options *options = nullptr,
                                                               bool throw_if_unsatisfied = false) {
    detail::same_length("ram_permutation_entry_point", w.unsorted_queue_witness.size(), w.unsorted_queue_prev_states.size());
    detail::same_length("ram_permutation_entry_point", w.sorted_queue_witness.size(), w.sorted_queue_prev_states.size());
    Result<zkc_ram_closed_form> r;
    r.closed_form_input = w.closed_form_input;
    uint64_t *trace = detail::make_trace(r, ZKC_RAM_NUM_COLS, limit, want_trace);
    const int rc = zkc_ram_permutation_entry_point(e.handle(), &r.closed_form_input, detail::ptr(w.unsorted_queue_witness),
                                                   detail::flat(w.unsorted_queue_prev_states), w.unsorted_queue_witness.size(),
                                                   detail::ptr(w.sorted_queue_witness), detail::flat(w.sorted_queue_prev_states),
                                                   w.sorted_queue_witness.size(), limit, options, 0, trace, r.commitment.data(), &r.status);
    detail::finish("ram_permutation_entry_point", rc, r, throw_if_unsatisfied);
    return r;
}

// FullStateCircuitQueue::push of a whole queue (how the reference tests build their inputs, ram_permutation/mod.rs:506-515)
inline zkc_queue_state12 memory_queue_simulate(Engine &e, const std::vector<zkc_memory_query> &records, std::vector<State12> &prev_states) {
    prev_states.assign(records.size(), State12{});
    zkc_queue_state12 fin{};
    const int rc = zkc_memory_queue_simulate(e.handle(), detail::ptr(records), records.size(), 1,
                                             records.empty() ? nullptr : prev_states.front().data(), &fin, 0);
    if (rc != ZKC_OK) throw Error("zkc_memory_queue_simulate", rc, zkc_status{rc, 0, -1, 0, 0});
    return fin;
}

// ---- log_sorter / storage_validity_by_grand_product ------------------------------------------------------------------
// EventsDeduplicatorInstanceWitness, /root/reference/src/log_sorter/input.rs:98-106
struct EventsDeduplicatorInstanceWitness {
    zkc_events_closed_form closed_form_input{};
    std::vector<zkc_log_query> initial_queue_witness;
    std::vector<State4> initial_queue_prev_tails;
    std::vector<zkc_log_query> intermediate_sorted_queue_witness;
    std::vector<State4> intermediate_sorted_queue_prev_tails;
    std::vector<State4> result_queue_tails;  // optional hint: tail after every executed push
};

// sort_and_deduplicate_events_entry_point, /root/reference/src/log_sorter/mod.rs:34-232
inline Result<zkc_events_closed_form> sort_and_deduplicate_events_entry_point(Engine &e, const EventsDeduplicatorInstanceWitness &w,
                                                                              size_t limit, bool want_trace = true,
                                                                              const zkc_sorter_options *options = nullptr,
                                                                              bool throw_if_unsatisfied = false) {
    detail::same_length("sort_and_deduplicate_events_entry_point", w.initial_queue_witness.size(), w.initial_queue_prev_tails.size());
    detail::same_length("sort_and_deduplicate_events_entry_point", w.intermediate_sorted_queue_witness.size(),
                        w.intermediate_sorted_queue_prev_tails.size());
    Result<zkc_events_closed_form> r;
    r.closed_form_input = w.closed_form_input;
    uint64_t *trace = detail::make_trace(r, ZKC_EV_NUM_COLS, limit, want_trace);
    const int rc = zkc_log_sorter_entry_point(e.handle(), &r.closed_form_input, detail::ptr(w.initial_queue_witness),
                                              detail::flat(w.initial_queue_prev_tails), w.initial_queue_witness.size(),
                                              detail::ptr(w.intermediate_sorted_queue_witness),
                                              detail::flat(w.intermediate_sorted_queue_prev_tails),
                                              w.intermediate_sorted_queue_witness.size(), detail::flat(w.result_queue_tails),
                                              w.result_queue_tails.size(), limit, options, 0, trace, r.commitment.data(), &r.status);
    detail::finish("sort_and_deduplicate_events_entry_point", rc, r, throw_if_unsatisfied);
    return r;
}

// StorageDeduplicatorInstanceWitness, /root/reference/src/storage_validity_by_grand_product/input.rs:128-136
struct StorageDeduplicatorInstanceWitness {
    zkc_storage_closed_form closed_form_input{};
    std::vector<zkc_log_query> unsorted_queue_witness;
    std::vector<State4> unsorted_queue_prev_tails;
    std::vector<zkc_log_query> intermediate_sorted_queue_witness;
    std::vector<uint32_t> intermediate_sorted_queue_timestamps;  // TimestampedStorageLogRecord.timestamp
    std::vector<State4> intermediate_sorted_queue_prev_tails;
    std::vector<State4> result_queue_tails;  // optional hint
};

// sort_and_deduplicate_storage_access_entry_point, /root/reference/src/storage_validity_by_grand_product/mod.rs:166-507
inline Result<zkc_storage_closed_form> sort_and_deduplicate_storage_access_entry_point(Engine &e, const StorageDeduplicatorInstanceWitness &w,
                                                                                       size_t limit, bool want_trace = true,
                                                                                       const zkc_sorter_options *options = nullptr,
                                                                                       bool throw_if_unsatisfied = false) {
    detail::same_length("sort_and_deduplicate_storage_access_entry_point", w.unsorted_queue_witness.size(), w.unsorted_queue_prev_tails.size());
    detail::same_length("sort_and_deduplicate_storage_access_entry_point", w.intermediate_sorted_queue_witness.size(),
                        w.intermediate_sorted_queue_prev_tails.size());
    detail::same_length("sort_and_deduplicate_storage_access_entry_point", w.intermediate_sorted_queue_witness.size(),
                        w.intermediate_sorted_queue_timestamps.size());
    Result<zkc_storage_closed_form> r;
    r.closed_form_input = w.closed_form_input;
    uint64_t *trace = detail::make_trace(r, ZKC_ST_NUM_COLS, limit, want_trace);
    const int rc = zkc_storage_validity_entry_point(
        e.handle(), &r.closed_form_input, detail::ptr(w.unsorted_queue_witness), detail::flat(w.unsorted_queue_prev_tails),
        w.unsorted_queue_witness.size(), detail::ptr(w.intermediate_sorted_queue_witness), detail::ptr(w.intermediate_sorted_queue_timestamps),
        detail::flat(w.intermediate_sorted_queue_prev_tails), w.intermediate_sorted_queue_witness.size(), detail::flat(w.result_queue_tails),
        w.result_queue_tails.size(), limit, options, 0, trace, r.commitment.data(), &r.status);
    detail::finish("sort_and_deduplicate_storage_access_entry_point", rc, r, throw_if_unsatisfied);
    return r;
}

// CircuitQueue::push of a whole LogQuery queue (log_sorter/mod.rs:571-580); extra_timestamps: storage sorter's sorted side
inline zkc_queue_state4 log_queue_simulate(Engine &e, const std::vector<zkc_log_query> &records, std::vector<State4> &prev_tails,
                                           const std::vector<uint32_t> *extra_timestamps = nullptr) {
    prev_tails.assign(records.size(), State4{});
    zkc_queue_state4 fin{};
    const int rc = zkc_log_queue_simulate(e.handle(), detail::ptr(records), extra_timestamps ? detail::ptr(*extra_timestamps) : nullptr,
                                          records.size(), 1, records.empty() ? nullptr : prev_tails.front().data(), &fin, 0);
    if (rc != ZKC_OK) throw Error("zkc_log_queue_simulate", rc, zkc_status{rc, 0, -1, 0, 0});
    return fin;
}

// ---- sort_decommittment_requests ---------------------------------------------------------------------------------------
// CodeDecommittmentsDeduplicatorInstanceWitness, /root/reference/src/sort_decommittment_requests/input.rs:114-131
struct CodeDecommittmentsDeduplicatorInstanceWitness {
    zkc_decommit_sorter_closed_form closed_form_input{};
    std::vector<zkc_decommit_query> initial_queue_witness;
    std::vector<State12> initial_queue_prev_states;
    std::vector<zkc_decommit_query> sorted_queue_witness;
    std::vector<State12> sorted_queue_prev_states;
    std::vector<State12> result_queue_states;  // optional hint: result-queue state after every executed push
};

// sort_and_deduplicate_code_decommittments_entry_point, /root/reference/src/sort_decommittment_requests/mod.rs:40-233
inline Result<zkc_decommit_sorter_closed_form> sort_and_deduplicate_code_decommittments_entry_point(
    Engine &e, const CodeDecommittmentsDeduplicatorInstanceWitness &w, size_t limit, bool want_trace = true,
    const zkc_sorter_options *options = nullptr, bool throw_if_unsatisfied = false) {
    detail::same_length("sort_and_deduplicate_code_decommittments_entry_point", w.initial_queue_witness.size(), w.initial_queue_prev_states.size());
    detail::same_length("sort_and_deduplicate_code_decommittments_entry_point", w.sorted_queue_witness.size(), w.sorted_queue_prev_states.size());
    Result<zkc_decommit_sorter_closed_form> r;
    r.closed_form_input = w.closed_form_input;
    uint64_t *trace = detail::make_trace(r, ZKC_DQ_NUM_COLS, limit, want_trace);
    const int rc = zkc_sort_decommittments_entry_point(
        e.handle(), &r.closed_form_input, detail::ptr(w.initial_queue_witness), detail::flat(w.initial_queue_prev_states),
        w.initial_queue_witness.size(), detail::ptr(w.sorted_queue_witness), detail::flat(w.sorted_queue_prev_states),
        w.sorted_queue_witness.size(), detail::flat(w.result_queue_states), w.result_queue_states.size(), limit, options, 0, trace,
        r.commitment.data(), &r.status);
    detail::finish("sort_and_deduplicate_code_decommittments_entry_point", rc, r, throw_if_unsatisfied);
    return r;
}

inline zkc_queue_state12 decommit_queue_simulate(Engine &e, const std::vector<zkc_decommit_query> &records, std::vector<State12> &prev_states) {
    prev_states.assign(records.size(), State12{});
    zkc_queue_state12 fin{};
    const int rc = zkc_decommit_queue_simulate(e.handle(), detail::ptr(records), records.size(), 1,
                                               records.empty() ? nullptr : prev_states.front().data(), &fin, 0);
    if (rc != ZKC_OK) throw Error("zkc_decommit_queue_simulate", rc, zkc_status{rc, 0, -1, 0, 0});
    return fin;
}

// ---- demux_log_queue ---------------------------------------------------------------------------------------------------
// LogDemuxerCircuitInstanceWitness, /root/reference/src/demux_log_queue/input.rs:124-129
struct LogDemuxerCircuitInstanceWitness {
    zkc_demux_closed_form closed_form_input{};
    std::vector<zkc_log_query> initial_queue_witness;
    std::vector<State4> initial_queue_prev_tails;
    // optional hint: per output queue (LogType order) its tail after each executed push
    std::array<std::vector<State4>, ZKC_DEMUX_NUM_QUEUES> output_queue_tails;
    bool have_output_queue_tails = false;
};

// demultiplex_storage_logs_enty_point (sic), /root/reference/src/demux_log_queue/mod.rs:38-217
inline Result<zkc_demux_closed_form> demultiplex_storage_logs_enty_point(Engine &e, const LogDemuxerCircuitInstanceWitness &w, size_t limit,
                                                                         bool want_trace = true, const zkc_demux_options *options = nullptr,
                                                                         bool throw_if_unsatisfied = false) {
    detail::same_length("demultiplex_storage_logs_enty_point", w.initial_queue_witness.size(), w.initial_queue_prev_tails.size());
    Result<zkc_demux_closed_form> r;
    r.closed_form_input = w.closed_form_input;
    uint64_t *trace = detail::make_trace(r, ZKC_DMX_NUM_COLS, limit, want_trace);
    std::vector<State4> tails;
    size_t counts[ZKC_DEMUX_NUM_QUEUES] = {0, 0, 0, 0, 0, 0};
    if (w.have_output_queue_tails)
        for (int q = 0; q < ZKC_DEMUX_NUM_QUEUES; q++) {
            counts[q] = w.output_queue_tails[q].size();
            tails.insert(tails.end(), w.output_queue_tails[q].begin(), w.output_queue_tails[q].end());
        }
    if (w.have_output_queue_tails && tails.empty()) tails.push_back(State4{});  // a non-null pointer says "hints supplied"
    const int rc = zkc_demux_log_queue_entry_point(e.handle(), &r.closed_form_input, detail::ptr(w.initial_queue_witness),
                                                   detail::flat(w.initial_queue_prev_tails), w.initial_queue_witness.size(),
                                                   w.have_output_queue_tails ? tails.front().data() : nullptr, counts, limit, options, 0,
                                                   trace, r.commitment.data(), &r.status);
    detail::finish("demultiplex_storage_logs_enty_point", rc, r, throw_if_unsatisfied);
    return r;
}

// ---- linear_hasher -----------------------------------------------------------------------------------------------------
// LinearHasherCircuitInstanceWitness, /root/reference/src/linear_hasher/input.rs:74-87
struct LinearHasherCircuitInstanceWitness {
    zkc_linear_hasher_closed_form closed_form_input{};
    std::vector<zkc_log_query> queue_witness;
    std::vector<State4> queue_prev_tails;
    // optional hint: the keccak state (25 lanes) after every cycle ([limit] entries); empty = rebuilt on the device
    std::vector<std::array<uint64_t, 25>> keccak_states;
};

// linear_hasher_entry_point, /root/reference/src/linear_hasher/mod.rs:35-214
inline Result<zkc_linear_hasher_closed_form> linear_hasher_entry_point(Engine &e, const LinearHasherCircuitInstanceWitness &w, size_t limit,
                                                                       bool want_trace = true, const zkc_sorter_options *options = nullptr,
                                                                       bool throw_if_unsatisfied = false) {
    detail::same_length("linear_hasher_entry_point", w.queue_witness.size(), w.queue_prev_tails.size());
    if (!w.keccak_states.empty() && w.keccak_states.size() < limit)
        throw Error("linear_hasher_entry_point: keccak_states holds the state after EVERY cycle", ZKC_ERR_INVALID_ARGUMENT,
                    zkc_status{ZKC_ERR_INVALID_ARGUMENT, 0, -1, 0, 0});
    Result<zkc_linear_hasher_closed_form> r;
    r.closed_form_input = w.closed_form_input;
    uint64_t *trace = detail::make_trace(r, ZKC_LH_NUM_COLS, limit, want_trace);
    const int rc = zkc_linear_hasher_entry_point(e.handle(), &r.closed_form_input, detail::ptr(w.queue_witness), detail::flat(w.queue_prev_tails),
                                                 w.queue_witness.size(), w.keccak_states.empty() ? nullptr : w.keccak_states.front().data(), limit,
                                                 options, 0, trace, r.commitment.data(), &r.status);
    detail::finish("linear_hasher_entry_point", rc, r, throw_if_unsatisfied);
    return r;
}

// ---- precompile round-function circuits ------------------------------------------------------------------------------------
// Keccak256RoundFunctionCircuitInstanceWitness (keccak256_round_function/input.rs:92-99) and its sha256 twin
// (sha256_round_function/input.rs): requests queue witness + the memory reads the circuit pops conditionally
template <class ClosedForm>
struct PrecompileCircuitInstanceWitness {
    ClosedForm closed_form_input{};
    std::vector<zkc_log_query> requests_queue_witness;
    std::vector<State4> requests_queue_prev_tails;
    std::vector<uint32_t> memory_reads_witness;       // 8 little-endian u32 limbs per 256-bit word, in pop order
    std::vector<State12> memory_queue_states;         // optional hint: memory-queue state after every executed push
};
using Keccak256RoundFunctionCircuitInstanceWitness = PrecompileCircuitInstanceWitness<zkc_keccak_closed_form>;
using Sha256RoundFunctionCircuitInstanceWitness = PrecompileCircuitInstanceWitness<zkc_sha256_closed_form>;

// keccak256_round_function_entry_point, /root/reference/src/keccak256_round_function/mod.rs:673-794
inline Result<zkc_keccak_closed_form> keccak256_round_function_entry_point(Engine &e, const Keccak256RoundFunctionCircuitInstanceWitness &w,
                                                                           size_t limit, bool want_trace = true,
                                                                           const zkc_precompile_options *options = nullptr,
                                                                           bool throw_if_unsatisfied = false) {
    detail::same_length("keccak256_round_function_entry_point", w.requests_queue_witness.size(), w.requests_queue_prev_tails.size());
    Result<zkc_keccak_closed_form> r;
    r.closed_form_input = w.closed_form_input;
    uint64_t *trace = detail::make_trace(r, ZKC_KC_NUM_COLS, limit, want_trace);
    const int rc = zkc_keccak256_round_function_entry_point(
        e.handle(), &r.closed_form_input, detail::ptr(w.requests_queue_witness), detail::flat(w.requests_queue_prev_tails),
        w.requests_queue_witness.size(), detail::ptr(w.memory_reads_witness), w.memory_reads_witness.size() / 8,
        detail::flat(w.memory_queue_states), w.memory_queue_states.size(), limit, options, 0, trace, r.commitment.data(), &r.status);
    detail::finish("keccak256_round_function_entry_point", rc, r, throw_if_unsatisfied);
    return r;
}

// sha256_round_function_entry_point, /root/reference/src/sha256_round_function/mod.rs:343-470
inline Result<zkc_sha256_closed_form> sha256_round_function_entry_point(Engine &e, const Sha256RoundFunctionCircuitInstanceWitness &w,
                                                                        size_t limit, bool want_trace = true,
                                                                        const zkc_precompile_options *options = nullptr,
                                                                        bool throw_if_unsatisfied = false) {
    detail::same_length("sha256_round_function_entry_point", w.requests_queue_witness.size(), w.requests_queue_prev_tails.size());
    Result<zkc_sha256_closed_form> r;
    r.closed_form_input = w.closed_form_input;
    uint64_t *trace = detail::make_trace(r, ZKC_SH_NUM_COLS, limit, want_trace);
    const int rc = zkc_sha256_round_function_entry_point(
        e.handle(), &r.closed_form_input, detail::ptr(w.requests_queue_witness), detail::flat(w.requests_queue_prev_tails),
        w.requests_queue_witness.size(), detail::ptr(w.memory_reads_witness), w.memory_reads_witness.size() / 8,
        detail::flat(w.memory_queue_states), w.memory_queue_states.size(), limit, options, 0, trace, r.commitment.data(), &r.status);
    detail::finish("sha256_round_function_entry_point", rc, r, throw_if_unsatisfied);
    return r;
}

// ---- code_unpacker_sha256 -----------------------------------------------------------------------------------------------------
// CodeDecommitterCircuitInstanceWitness, /root/reference/src/code_unpacker_sha256/input.rs:152-160
struct CodeDecommitterCircuitInstanceWitness {
    zkc_code_unpacker_closed_form closed_form_input{};
    std::vector<zkc_decommit_query> sorted_requests_queue_witness;
    std::vector<State12> sorted_requests_queue_prev_states;
    std::vector<std::array<uint32_t, 8>> code_words;  // flattened over the requests, little-endian u32 limbs per 256-bit word
    std::vector<State12> memory_queue_states;         // optional hint: memory-queue state after every executed write
};

// unpack_code_into_memory_entry_point, /root/reference/src/code_unpacker_sha256/mod.rs:33-148
inline Result<zkc_code_unpacker_closed_form> unpack_code_into_memory_entry_point(Engine &e, const CodeDecommitterCircuitInstanceWitness &w,
                                                                                 size_t limit, bool want_trace = true,
                                                                                 const zkc_sorter_options *options = nullptr,
                                                                                 bool throw_if_unsatisfied = false) {
    detail::same_length("unpack_code_into_memory_entry_point", w.sorted_requests_queue_witness.size(), w.sorted_requests_queue_prev_states.size());
    Result<zkc_code_unpacker_closed_form> r;
    r.closed_form_input = w.closed_form_input;
    uint64_t *trace = detail::make_trace(r, ZKC_CU_NUM_COLS, limit, want_trace);
    const int rc = zkc_code_unpacker_entry_point(
        e.handle(), &r.closed_form_input, detail::ptr(w.sorted_requests_queue_witness), detail::flat(w.sorted_requests_queue_prev_states),
        w.sorted_requests_queue_witness.size(), w.code_words.empty() ? nullptr : w.code_words.front().data(), w.code_words.size(),
        detail::flat(w.memory_queue_states), w.memory_queue_states.size(), limit, options, 0, trace, r.commitment.data(), &r.status);
    detail::finish("unpack_code_into_memory_entry_point", rc, r, throw_if_unsatisfied);
    return r;
}

// ---- main_vm ---------------------------------------------------------------------------------------------------------------
// VmCircuitWitness, /root/reference/src/fsm_input_output/circuit_inputs/main_vm.rs:64-71, with the WitnessOracle flattened
// per cycle: `snapshots` = the VmLocalState before every cycle + the final one (limit + 1),
// `witness_oracle` = one record of oracle answers per cycle, `callstack_witness` = the frames rets pop, indexed from the
// per-cycle record
struct VmCircuitWitness {
    zkc_vm_closed_form closed_form_input{};
    const zkc_vm_isa *isa = nullptr;  // zkevm_opcode_defs tables (un-vendored: supplied by the caller)
    std::vector<zkc_vm_state> snapshots;
    std::vector<zkc_vm_cycle_witness> witness_oracle;
    std::vector<zkc_vm_callstack_witness> callstack_witness;
};

// main_vm_entry_point, /root/reference/src/main_vm/mod.rs:47-236
inline Result<zkc_vm_closed_form> main_vm_entry_point(Engine &e, const VmCircuitWitness &w, size_t limit, bool want_trace = true,
                                                      const zkc_vm_options *options = nullptr, bool throw_if_unsatisfied = false) {
    if (!w.isa) throw std::invalid_argument("main_vm_entry_point: no ISA tables");
    if (w.snapshots.size() < limit + 1 || w.witness_oracle.size() < limit)
        throw std::invalid_argument("main_vm_entry_point: limit + 1 snapshots and limit oracle records are required");
    Result<zkc_vm_closed_form> r;
    r.closed_form_input = w.closed_form_input;
    uint64_t *trace = detail::make_trace(r, ZKC_VM_NUM_COLS, limit, want_trace);
    const int rc = zkc_main_vm_entry_point(e.handle(), &r.closed_form_input, w.isa, w.snapshots.data(), detail::ptr(w.witness_oracle),
                                           detail::ptr(w.callstack_witness), w.callstack_witness.size(), limit, options, 0, trace,
                                           r.commitment.data(), &r.status);
    detail::finish("main_vm_entry_point", rc, r, throw_if_unsatisfied);
    return r;
}

// initial_bootloader_state, /root/reference/src/main_vm/loading.rs:13-226
inline zkc_vm_state main_vm_initial_state(Engine &e, const zkc_vm_closed_form &io, const zkc_vm_isa &isa) {
    zkc_vm_state st;
    std::memset(&st, 0, sizeof st);
    const int rc = zkc_main_vm_initial_state(e.handle(), &io, &isa, &st);
    if (rc != ZKC_OK) throw Error("zkc_main_vm_initial_state", rc, zkc_status{rc, 0, -1, 0, 0});
    return st;
}

// the cells the add/sub, binop, mul/div and shift gadgets allocate on every cycle + the per-cycle relations (zkc_b200.h,
// ZKC_VM_GADGET_COLUMNS), from a finished DENSE trace [ZKC_VM_NUM_COLS][limit] (host)
inline std::vector<uint64_t> main_vm_gadget_cells(Engine &e, const std::vector<uint64_t> &trace, size_t limit) {
    std::vector<uint64_t> out((size_t)ZKC_VMG_NUM_COLS * limit);
    const int rc = zkc_main_vm_gadget_cells(e.handle(), trace.data(), limit, 1, 0, out.data());
    if (rc != ZKC_OK) throw Error("zkc_main_vm_gadget_cells", rc, zkc_status{rc, 0, -1, 0, 0});
    return out;
}

// the cells the ptr, jump and context gadgets allocate on every cycle (zkc_b200.h, ZKC_VM_STATE_GADGET_COLUMNS), from a finished DENSE
// trace [ZKC_VM_NUM_COLS][limit] and the [limit + 1] snapshots the entry point took (host)
inline std::vector<uint64_t> main_vm_state_gadget_cells(Engine &e, const std::vector<uint64_t> &trace, const std::vector<zkc_vm_state> &snapshots,
                                                        size_t limit) {
    if (trace.size() < (size_t)ZKC_VM_NUM_COLS * limit || snapshots.size() < limit + 1)
        throw Error("main_vm_state_gadget_cells", ZKC_ERR_INVALID_ARGUMENT, zkc_status{ZKC_ERR_INVALID_ARGUMENT, 0, -1, 0, 0});
    std::vector<uint64_t> out((size_t)ZKC_VMS_NUM_COLS * limit);
    const int rc = zkc_main_vm_state_gadget_cells(e.handle(), trace.data(), snapshots.data(), limit, 1, 0, out.data());
    if (rc != ZKC_OK) throw Error("zkc_main_vm_state_gadget_cells", rc, zkc_status{rc, 0, -1, 0, 0});
    return out;
}

// the three memory-queue relations every cycle evaluates whatever its opcode -- opcode fetch, src0 read, dst0 write: initial state,
// permutation output, selected tail and length per step (zkc_b200.h, ZKC_VM_MEMORY_SPONGE_COLUMNS); same inputs as above (host)
inline std::vector<uint64_t> main_vm_memory_sponge_cells(Engine &e, const std::vector<uint64_t> &trace, const std::vector<zkc_vm_state> &snapshots,
                                                         size_t limit) {
    if (trace.size() < (size_t)ZKC_VM_NUM_COLS * limit || snapshots.size() < limit + 1)
        throw Error("main_vm_memory_sponge_cells", ZKC_ERR_INVALID_ARGUMENT, zkc_status{ZKC_ERR_INVALID_ARGUMENT, 0, -1, 0, 0});
    std::vector<uint64_t> out((size_t)ZKC_VMQ_NUM_COLS * limit);
    const int rc = zkc_main_vm_memory_sponge_cells(e.handle(), trace.data(), snapshots.data(), limit, 1, 0, out.data());
    if (rc != ZKC_OK) throw Error("zkc_main_vm_memory_sponge_cells", rc, zkc_status{rc, 0, -1, 0, 0});
    return out;
}

// the cells create_prestate allocates on the way to the values the DENSE trace names -- selector masks, the 15-step register select
// chains, operand locations, src0 selects, swap, erasure flags (zkc_b200.h, ZKC_VM_PRESTATE_COLUMNS); same inputs as above (host)
inline std::vector<uint64_t> main_vm_prestate_cells(Engine &e, const std::vector<uint64_t> &trace, const std::vector<zkc_vm_state> &snapshots,
                                                    size_t limit) {
    if (trace.size() < (size_t)ZKC_VM_NUM_COLS * limit || snapshots.size() < limit + 1)
        throw Error("main_vm_prestate_cells", ZKC_ERR_INVALID_ARGUMENT, zkc_status{ZKC_ERR_INVALID_ARGUMENT, 0, -1, 0, 0});
    std::vector<uint64_t> out((size_t)ZKC_VMP_NUM_COLS * limit);
    const int rc = zkc_main_vm_prestate_cells(e.handle(), trace.data(), snapshots.data(), limit, 1, 0, out.data());
    if (rc != ZKC_OK) throw Error("zkc_main_vm_prestate_cells", rc, zkc_status{rc, 0, -1, 0, 0});
    return out;
}

// the register write-back of the state diffs (cycle.rs:158-433): dst0 / dst1 update flags, the far call / far return register conventions,
// both is_pointer dot products and the value select chain of each of the 15 registers (zkc_b200.h, ZKC_VM_WRITEBACK_COLUMNS) (host)
inline std::vector<uint64_t> main_vm_writeback_cells(Engine &e, const zkc_vm_isa &isa, const std::vector<uint64_t> &trace,
                                                     const std::vector<zkc_vm_state> &snapshots, size_t limit) {
    if (trace.size() < (size_t)ZKC_VM_NUM_COLS * limit || snapshots.size() < limit + 1)
        throw Error("main_vm_writeback_cells", ZKC_ERR_INVALID_ARGUMENT, zkc_status{ZKC_ERR_INVALID_ARGUMENT, 0, -1, 0, 0});
    std::vector<uint64_t> out((size_t)ZKC_VMW_NUM_COLS * limit);
    const int rc = zkc_main_vm_writeback_cells(e.handle(), &isa, trace.data(), snapshots.data(), limit, 1, 0, out.data());
    if (rc != ZKC_OK) throw Error("zkc_main_vm_writeback_cells", rc, zkc_status{rc, 0, -1, 0, 0});
    return out;
}

// constraint evaluation of finished traces (host buffers): violating rows; st describes the first one
inline uint64_t ram_permutation_check_trace(Engine &e, const zkc_ram_closed_form &io, const std::vector<uint64_t> &trace, size_t limit,
                                            uint32_t gates = 0, zkc_status *st = nullptr) {
    uint64_t v = 0;
    zkc_status local;
    const int rc = zkc_ram_permutation_check_trace(e.handle(), &io, trace.data(), limit, nullptr, gates, 0, &v, st ? st : &local);
    if (rc != ZKC_OK && rc != ZKC_ERR_UNSATISFIED) throw Error("zkc_ram_permutation_check_trace", rc, st ? *st : local);
    return v;
}
inline uint64_t log_sorter_check_trace(Engine &e, const zkc_events_closed_form &io, const std::vector<uint64_t> &trace, size_t limit,
                                       uint32_t gates = 0, zkc_status *st = nullptr) {
    uint64_t v = 0;
    zkc_status local;
    const int rc = zkc_log_sorter_check_trace(e.handle(), &io, trace.data(), limit, gates, 0, &v, st ? st : &local);
    if (rc != ZKC_OK && rc != ZKC_ERR_UNSATISFIED) throw Error("zkc_log_sorter_check_trace", rc, st ? *st : local);
    return v;
}
inline uint64_t storage_validity_check_trace(Engine &e, const zkc_storage_closed_form &io, const std::vector<uint64_t> &trace, size_t limit,
                                             uint32_t gates = 0, zkc_status *st = nullptr) {
    uint64_t v = 0;
    zkc_status local;
    const int rc = zkc_storage_validity_check_trace(e.handle(), &io, trace.data(), limit, gates, 0, &v, st ? st : &local);
    if (rc != ZKC_OK && rc != ZKC_ERR_UNSATISFIED) throw Error("zkc_storage_validity_check_trace", rc, st ? *st : local);
    return v;
}
inline uint64_t sort_decommittments_check_trace(Engine &e, const zkc_decommit_sorter_closed_form &io, const std::vector<uint64_t> &trace, size_t limit,
                                                uint32_t gates = 0, zkc_status *st = nullptr) {
    uint64_t v = 0;
    zkc_status local;
    const int rc = zkc_sort_decommittments_check_trace(e.handle(), &io, trace.data(), limit, gates, 0, &v, st ? st : &local);
    if (rc != ZKC_OK && rc != ZKC_ERR_UNSATISFIED) throw Error("zkc_sort_decommittments_check_trace", rc, st ? *st : local);
    return v;
}
inline uint64_t demux_log_queue_check_trace(Engine &e, const zkc_demux_closed_form &io, const std::vector<uint64_t> &trace, size_t limit,
                                            uint32_t gates = 0, const zkc_demux_options *options = nullptr, zkc_status *st = nullptr) {
    uint64_t v = 0;
    zkc_status local;
    const int rc = zkc_demux_log_queue_check_trace(e.handle(), &io, options, trace.data(), limit, gates, 0, &v, st ? st : &local);
    if (rc != ZKC_OK && rc != ZKC_ERR_UNSATISFIED) throw Error("zkc_demux_log_queue_check_trace", rc, st ? *st : local);
    return v;
}
inline uint64_t sha256_round_function_check_trace(Engine &e, const zkc_sha256_closed_form &io, const std::vector<uint64_t> &trace, size_t limit,
                                                  uint32_t gates = 0, const zkc_precompile_options *options = nullptr, zkc_status *st = nullptr) {
    uint64_t v = 0;
    zkc_status local;
    const int rc = zkc_sha256_round_function_check_trace(e.handle(), &io, options, trace.data(), limit, gates, 0, &v, st ? st : &local);
    if (rc != ZKC_OK && rc != ZKC_ERR_UNSATISFIED) throw Error("zkc_sha256_round_function_check_trace", rc, st ? *st : local);
    return v;
}
inline uint64_t code_unpacker_check_trace(Engine &e, const zkc_code_unpacker_closed_form &io, const std::vector<uint64_t> &trace, size_t limit,
                                          uint32_t gates = 0, zkc_status *st = nullptr) {
    uint64_t v = 0;
    zkc_status local;
    const int rc = zkc_code_unpacker_check_trace(e.handle(), &io, trace.data(), limit, gates, 0, &v, st ? st : &local);
    if (rc != ZKC_OK && rc != ZKC_ERR_UNSATISFIED) throw Error("zkc_code_unpacker_check_trace", rc, st ? *st : local);
    return v;
}
inline uint64_t linear_hasher_check_trace(Engine &e, const zkc_linear_hasher_closed_form &io, const std::vector<uint64_t> &trace, size_t limit,
                                          uint32_t gates = 0, zkc_status *st = nullptr) {
    uint64_t v = 0;
    zkc_status local;
    const int rc = zkc_linear_hasher_check_trace(e.handle(), &io, trace.data(), limit, gates, 0, &v, st ? st : &local);
    if (rc != ZKC_OK && rc != ZKC_ERR_UNSATISFIED) throw Error("zkc_linear_hasher_check_trace", rc, st ? *st : local);
    return v;
}
inline uint64_t keccak256_round_function_check_trace(Engine &e, const zkc_keccak_closed_form &io, const std::vector<uint64_t> &trace, size_t limit,
                                                     uint32_t gates = 0, const zkc_precompile_options *options = nullptr, zkc_status *st = nullptr) {
    uint64_t v = 0;
    zkc_status local;
    const int rc = zkc_keccak256_round_function_check_trace(e.handle(), &io, options, trace.data(), limit, gates, 0, &v, st ? st : &local);
    if (rc != ZKC_OK && rc != ZKC_ERR_UNSATISFIED) throw Error("zkc_keccak256_round_function_check_trace", rc, st ? *st : local);
    return v;
}

}  // namespace zkc_b200
