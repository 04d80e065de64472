// C++ host mirror (include/zkc_b200.hpp) against the CPU oracle, written the way the reference's own tests are
// (/root/reference/src/ram_permutation/mod.rs:395-557, sort_decommittment_requests/mod.rs:420-563, demux_log_queue/mod.rs:482-600):
// push the inputs into fresh queues, run the entry point, check the outcome.  TEST INFRASTRUCTURE: links oracle/liborc.so as
// the checker.
//   host_mirror_test nodevice   no GPU required: the engine must refuse to start (no CPU fallback) and say why
//   host_mirror_test parity     on a GPU: ram_permutation, sort_decommittment_requests, demux_log_queue, code_unpacker_sha256, linear_hasher
//                               bit-exact vs the oracle
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <numeric>

#include "zkc_b200.hpp"
extern "C" {
#include "oracle.h"
}

using namespace zkc_b200;

static uint64_t sm64(uint64_t &s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static int failures = 0;
#define CHECK(cond)                                                                  \
    do {                                                                             \
        if (!(cond)) { std::printf("FAILED %s:%d  %s\n", __FILE__, __LINE__, #cond); failures++; } \
    } while (0)

template <class T>
static bool same_bytes(const T &a, const T &b) { return std::memcmp(&a, &b, sizeof(T)) == 0; }

static void ram_parity(Engine &e) {
    const size_t n = 1000, limit = 1024;
    uint64_t seed = 0xC1;
    std::vector<zkc_memory_query> u(n);
    uint32_t cur[64][8] = {};
    for (size_t i = 0; i < n; i++) {
        zkc_memory_query q;
        std::memset(&q, 0, sizeof q);
        const uint32_t cell = (uint32_t)(sm64(seed) % 64);
        q.timestamp = 1000 + 4 * (uint32_t)i; q.memory_page = 20 + cell / 16; q.index = cell % 16;
        q.rw_flag = sm64(seed) % 100 < 60;
        if (q.rw_flag) for (int k = 0; k < 8; k++) cur[cell][k] = (uint32_t)sm64(seed);
        std::memcpy(q.value, cur[cell], 32);
        u[i] = q;
    }
    std::vector<zkc_memory_query> s = u;
    std::stable_sort(s.begin(), s.end(), [](const zkc_memory_query &a, const zkc_memory_query &b) {
        if (a.memory_page != b.memory_page) return a.memory_page < b.memory_page;
        if (a.index != b.index) return a.index < b.index;
        return a.timestamp < b.timestamp;
    });
    RamPermutationCircuitInstanceWitness w;
    w.unsorted_queue_witness = u; w.sorted_queue_witness = s;
    w.closed_form_input.start_flag = 1;
    w.closed_form_input.observable_input.unsorted_queue_initial_state = memory_queue_simulate(e, u, w.unsorted_queue_prev_states);
    w.closed_form_input.observable_input.sorted_queue_initial_state = memory_queue_simulate(e, s, w.sorted_queue_prev_states);
    // the engine's queue simulation against the oracle's
    std::vector<uint64_t> prev(12 * n);
    zkc_queue_state12 fin;
    orc_memory_queue_simulate(u.data(), n, prev.data(), &fin);
    CHECK(same_bytes(fin, w.closed_form_input.observable_input.unsorted_queue_initial_state));
    CHECK(std::memcmp(prev.data(), w.unsorted_queue_prev_states.data(), prev.size() * 8) == 0);

    const auto got = ram_permutation_entry_point(e, w, limit);
    zkc_ram_closed_form io = w.closed_form_input;
    std::vector<uint64_t> trace((size_t)ZKC_RAM_NUM_COLS * limit);
    uint64_t com[4];
    zkc_status st;
    const int rc = orc_ram_permutation_entry_point(&io, u.data(), n, s.data(), n, limit, nullptr, trace.data(), com, &st);
    CHECK(rc == ZKC_OK && got.status.code == ZKC_OK);
    CHECK(std::memcmp(com, got.commitment.data(), 32) == 0);
    CHECK(same_bytes(io.hidden_fsm_output, got.closed_form_input.hidden_fsm_output));
    CHECK(io.completion_flag == 1 && got.closed_form_input.completion_flag == 1);
    CHECK(std::memcmp(trace.data(), got.trace.data(), trace.size() * 8) == 0);
    // an unsatisfiable witness is a status, a broken queue witness a panic
    auto bad = w;
    std::swap(bad.sorted_queue_witness[10], bad.sorted_queue_witness[500]);
    bool threw = false;
    try { ram_permutation_entry_point(e, bad, limit); } catch (const Error &err) { threw = err.code == ZKC_ERR_QUEUE_WITNESS_INCONSISTENT; }
    CHECK(threw);
    std::printf("ram_permutation: %zu queries, commitment %016llx, launches so far %llu\n", n, (unsigned long long)got.commitment[0],
                (unsigned long long)e.launch_count());
}

static void decommit_parity(Engine &e) {
    const size_t n = 500, limit = 512, hashes = 20;
    uint64_t seed = 0xF1;
    uint32_t hash[hashes][8];
    for (auto &h : hash) { for (auto &l : h) l = (uint32_t)sm64(seed); h[7] |= 1; }
    std::vector<zkc_decommit_query> u(n);
    size_t first[hashes];
    std::fill(first, first + hashes, n);
    for (size_t i = 0; i < n; i++) {
        const size_t k = sm64(seed) % hashes;
        zkc_decommit_query q;
        std::memset(&q, 0, sizeof q);
        std::memcpy(q.code_hash, hash[k], 32);
        if (first[k] == n) first[k] = i;
        q.is_first = first[k] == i; q.page = 2048 + 8 * (uint32_t)first[k]; q.timestamp = 1000 + 4 * (uint32_t)i;
        u[i] = q;
    }
    std::vector<zkc_decommit_query> s = u;
    std::stable_sort(s.begin(), s.end(), [](const zkc_decommit_query &a, const zkc_decommit_query &b) {
        for (int l = 7; l >= 0; l--) if (a.code_hash[l] != b.code_hash[l]) return a.code_hash[l] < b.code_hash[l];
        return a.timestamp < b.timestamp;
    });
    CodeDecommittmentsDeduplicatorInstanceWitness w;
    w.initial_queue_witness = u; w.sorted_queue_witness = s;
    w.closed_form_input.start_flag = 1;
    w.closed_form_input.initial_queue_state = decommit_queue_simulate(e, u, w.initial_queue_prev_states);
    w.closed_form_input.sorted_queue_initial_state = decommit_queue_simulate(e, s, w.sorted_queue_prev_states);
    zkc_decommit_sorter_closed_form io = w.closed_form_input;
    std::vector<uint64_t> trace((size_t)ZKC_DQ_NUM_COLS * limit), states(12 * (limit + 1));
    size_t n_states = 0;
    uint64_t com[4];
    zkc_status st;
    const int rc = orc_sort_decommittments_entry_point(&io, u.data(), n, s.data(), n, limit, nullptr, trace.data(), states.data(), &n_states, com, &st);
    CHECK(rc == ZKC_OK && n_states == hashes && io.final_queue_state.length == hashes);
    for (int pass = 0; pass < 2; pass++) {  // without and with the result-queue hints
        if (pass) {
            w.result_queue_states.resize(n_states);
            std::memcpy(w.result_queue_states.data(), states.data(), n_states * 96);
        }
        const auto got = sort_and_deduplicate_code_decommittments_entry_point(e, w, limit);
        CHECK(got.status.code == ZKC_OK);
        CHECK(std::memcmp(com, got.commitment.data(), 32) == 0);
        CHECK(same_bytes(io.hidden_fsm_output, got.closed_form_input.hidden_fsm_output));
        CHECK(same_bytes(io.final_queue_state, got.closed_form_input.final_queue_state));
        CHECK(std::memcmp(trace.data(), got.trace.data(), trace.size() * 8) == 0);
    }
    std::printf("sort_decommittment_requests: %zu requests -> %zu records\n", n, n_states);
}

static void demux_parity(Engine &e) {
    const size_t n = 600, limit = 640;
    uint64_t seed = 0xD3;
    std::vector<zkc_log_query> recs(n);
    for (size_t i = 0; i < n; i++) {
        zkc_log_query q;
        std::memset(&q, 0, sizeof q);
        for (auto &l : q.address) l = (uint32_t)sm64(seed);
        for (auto &l : q.key) l = (uint32_t)sm64(seed);
        for (auto &l : q.read_value) l = (uint32_t)sm64(seed);
        for (auto &l : q.written_value) l = (uint32_t)sm64(seed);
        q.timestamp = 1000 + 4 * (uint32_t)i; q.tx_number_in_block = (uint32_t)(sm64(seed) % 1000);
        const uint32_t r = (uint32_t)(sm64(seed) % 100);
        const uint32_t kind = r < 50 ? 0 : r < 75 ? 1 : r < 85 ? 2 : r < 91 ? 3 : r < 97 ? 4 : 5;
        const uint32_t aux = kind < 3 ? kind : 3;
        q.flags = ZKC_LQ_FLAGS(aux, 0, 1, 0, 0);
        if (kind >= 3) {
            std::memset(q.address, 0, sizeof q.address);
            q.address[0] = kind == 3 ? 0x8010u : kind == 4 ? 0x02u : 0x01u;
        }
        recs[i] = q;
    }
    LogDemuxerCircuitInstanceWitness w;
    w.initial_queue_witness = recs;
    w.closed_form_input.start_flag = 1;
    w.closed_form_input.initial_log_queue_state = log_queue_simulate(e, recs, w.initial_queue_prev_tails);
    zkc_demux_closed_form io = w.closed_form_input;
    std::vector<uint64_t> trace((size_t)ZKC_DMX_NUM_COLS * limit), tails((size_t)6 * limit * 4);
    size_t counts[6];
    uint64_t com[4];
    zkc_status st;
    const int rc = orc_demux_log_queue_entry_point(&io, recs.data(), n, limit, nullptr, trace.data(), tails.data(), counts, com, &st);
    CHECK(rc == ZKC_OK && std::accumulate(counts, counts + 6, (size_t)0) == n);
    for (int pass = 0; pass < 2; pass++) {
        if (pass) {
            w.have_output_queue_tails = true;
            for (int q = 0; q < 6; q++) {
                w.output_queue_tails[q].resize(counts[q]);
                std::memcpy(w.output_queue_tails[q].data(), tails.data() + (size_t)q * limit * 4, counts[q] * 32);
            }
        }
        const auto got = demultiplex_storage_logs_enty_point(e, w, limit);
        CHECK(got.status.code == ZKC_OK);
        CHECK(std::memcmp(com, got.commitment.data(), 32) == 0);
        CHECK(same_bytes(io.hidden_fsm_output, got.closed_form_input.hidden_fsm_output));
        CHECK(std::memcmp(io.output_queue_states, got.closed_form_input.output_queue_states, sizeof io.output_queue_states) == 0);
        CHECK(std::memcmp(trace.data(), got.trace.data(), trace.size() * 8) == 0);
    }
    std::printf("demux_log_queue: %zu records -> %zu / %zu / %zu / %zu / %zu / %zu\n", n, counts[0], counts[1], counts[2], counts[3],
                counts[4], counts[5]);
}

static void linear_hasher_parity(Engine &e) {
    const size_t n = 300, limit = 320;
    uint64_t seed = 0x1A;
    std::vector<zkc_log_query> recs(n);
    for (size_t i = 0; i < n; i++) {
        zkc_log_query &q = recs[i];
        std::memset(&q, 0, sizeof q);
        for (auto &l : q.address) l = (uint32_t)sm64(seed);
        for (auto &l : q.key) l = (uint32_t)sm64(seed);
        for (auto &l : q.written_value) l = (uint32_t)sm64(seed);
        q.tx_number_in_block = (uint32_t)(sm64(seed) & 0xFFFF); q.timestamp = (uint32_t)i + 1;
        q.flags = ZKC_LQ_FLAGS(2, 0, 1, 0, (uint32_t)(sm64(seed) & 1));
    }
    LinearHasherCircuitInstanceWitness w;
    w.queue_witness = recs;
    w.closed_form_input.start_flag = 1;
    w.closed_form_input.queue_state = log_queue_simulate(e, recs, w.queue_prev_tails);
    zkc_linear_hasher_closed_form io = w.closed_form_input;
    std::vector<uint64_t> trace((size_t)ZKC_LH_NUM_COLS * limit), states(25 * limit);
    uint64_t com[4];
    zkc_status st;
    const int rc = orc_linear_hasher_entry_point(&io, recs.data(), n, limit, nullptr, trace.data(), states.data(), com, &st);
    CHECK(rc == ZKC_OK && io.completion_flag == 1);
    // the digest is Keccak-256 of the concatenated serialisations
    std::vector<uint8_t> stream(n * ZKC_LH_MESSAGE_BYTES);
    for (size_t i = 0; i < n; i++) orc_log_query_into_bytes(&recs[i], stream.data() + i * ZKC_LH_MESSAGE_BYTES);
    uint8_t digest[32];
    orc_keccak256(stream.data(), stream.size(), digest);
    for (int i = 0; i < 32; i++) CHECK(io.keccak256_hash[i] == digest[i]);
    for (int pass = 0; pass < 2; pass++) {
        if (pass) {
            w.keccak_states.resize(limit);
            std::memcpy(w.keccak_states.front().data(), states.data(), states.size() * 8);
        }
        const auto got = linear_hasher_entry_point(e, w, limit);
        CHECK(got.status.code == ZKC_OK);
        CHECK(std::memcmp(com, got.commitment.data(), 32) == 0);
        CHECK(std::memcmp(io.keccak256_hash, got.closed_form_input.keccak256_hash, sizeof io.keccak256_hash) == 0);
        CHECK(std::memcmp(trace.data(), got.trace.data(), trace.size() * 8) == 0);
    }
    std::printf("linear_hasher: %zu messages, %zu cycles\n", n, limit);
}

static void code_unpacker_parity(Engine &e) {
    // three bytecodes of 5, 1 and 9 words; versioned hash = SHA-256 of the code with the top 4 bytes replaced (mod.rs:187-213)
    const size_t lens[3] = {5, 1, 9};
    uint64_t seed = 0xCD;
    CodeDecommitterCircuitInstanceWitness w;
    size_t rounds = 0;
    for (size_t k = 0; k < 3; k++) {
        std::vector<uint32_t> be;  // the code as big-endian u32 words, the way SHA-256 consumes it
        for (size_t i = 0; i < lens[k]; i++) {
            std::array<uint32_t, 8> limbs;
            for (auto &l : limbs) l = (uint32_t)sm64(seed);
            w.code_words.push_back(limbs);
            for (int j = 7; j >= 0; j--) be.push_back(limbs[j]);
        }
        be.push_back(0x80000000u);
        while (be.size() % 16 != 15) be.push_back(0);
        be.push_back((uint32_t)(lens[k] * 256));
        uint32_t st[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
        for (size_t b = 0; b < be.size(); b += 16) orc_sha256_compress(st, be.data() + b);
        zkc_decommit_query q;
        std::memset(&q, 0, sizeof q);
        for (int i = 0; i < 7; i++) q.code_hash[i] = st[7 - i];
        q.code_hash[7] = (ZKC_CODE_HASH_VERSION_TOP16 << 16) | (uint32_t)lens[k];
        q.page = 2048 + 8 * (uint32_t)k; q.timestamp = 77 + (uint32_t)k; q.is_first = 1;
        w.sorted_requests_queue_witness.push_back(q);
        rounds += (lens[k] + 1) / 2;
    }
    const size_t limit = rounds + 3;
    w.closed_form_input.start_flag = 1;
    w.closed_form_input.sorted_requests_queue_initial_state =
        decommit_queue_simulate(e, w.sorted_requests_queue_witness, w.sorted_requests_queue_prev_states);
    zkc_code_unpacker_closed_form io = w.closed_form_input;
    std::vector<uint64_t> trace((size_t)ZKC_CU_NUM_COLS * limit), states(12 * (2 * limit + 1));
    size_t n_states = 0;
    uint64_t com[4];
    zkc_status st;
    const int rc = orc_code_unpacker_entry_point(&io, w.sorted_requests_queue_witness.data(), 3, w.code_words.front().data(),
                                                 w.code_words.size(), limit, nullptr, trace.data(), states.data(), &n_states, com, &st);
    CHECK(rc == ZKC_OK && io.completion_flag == 1 && n_states == 15);
    const auto got = unpack_code_into_memory_entry_point(e, w, limit);
    CHECK(got.status.code == ZKC_OK);
    CHECK(std::memcmp(com, got.commitment.data(), 32) == 0);
    CHECK(same_bytes(io.hidden_fsm_output, got.closed_form_input.hidden_fsm_output));
    CHECK(same_bytes(io.memory_queue_final_state, got.closed_form_input.memory_queue_final_state));
    CHECK(std::memcmp(trace.data(), got.trace.data(), trace.size() * 8) == 0);
    std::printf("code_unpacker_sha256: 3 bytecodes, %zu words, %zu cycles\n", w.code_words.size(), limit);
}

int main(int argc, char **argv) {
    const std::string mode = argc > 1 ? argv[1] : "nodevice";
    std::printf("%s\n", Engine::version().c_str());
    // every entry point of the mirror instantiates (the header is header-only: this is its compile check)
    void *instantiated[] = {(void *)&ram_permutation_entry_point, (void *)&sort_and_deduplicate_events_entry_point,
                            (void *)&sort_and_deduplicate_storage_access_entry_point,
                            (void *)&sort_and_deduplicate_code_decommittments_entry_point, (void *)&demultiplex_storage_logs_enty_point,
                            (void *)&unpack_code_into_memory_entry_point, (void *)&keccak256_round_function_entry_point, (void *)&sha256_round_function_entry_point,
                            (void *)&main_vm_entry_point, (void *)&main_vm_initial_state, (void *)&linear_hasher_entry_point,
                            (void *)&main_vm_gadget_cells, (void *)&main_vm_state_gadget_cells, (void *)&main_vm_memory_sponge_cells, (void *)&main_vm_prestate_cells, (void *)&main_vm_writeback_cells,
                            (void *)&ram_permutation_check_trace, (void *)&log_sorter_check_trace,
                            (void *)&storage_validity_check_trace, (void *)&sort_decommittments_check_trace,
                            (void *)&demux_log_queue_check_trace, (void *)&sha256_round_function_check_trace,
                            (void *)&code_unpacker_check_trace, (void *)&linear_hasher_check_trace,
                            (void *)&keccak256_round_function_check_trace};
    std::printf("%zu entry points\n", sizeof instantiated / sizeof instantiated[0]);
    if (mode == "nodevice") {
        try {
            Engine e(0);
            std::printf("a device is present: nothing to check in this mode\n");
        } catch (const Error &err) {
            if (err.code != ZKC_ERR_NO_DEVICE) { std::printf("unexpected error: %s\n", err.what()); return 1; }
            std::printf("refused without a device: %s\n", err.what());
        }
        return 0;
    }
    Engine e(0);
    ram_parity(e);
    decommit_parity(e);
    demux_parity(e);
    code_unpacker_parity(e);
    linear_hasher_parity(e);
    std::printf(failures ? "%d check(s) FAILED\n" : "all checks passed\n", failures);
    return failures ? 1 : 0;
}
