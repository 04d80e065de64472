"""Synthetic, *valid* circuit inputs.  Host-side input preparation only (numpy);
in production these arrays come from the out-of-circuit VM run.  Values are built by constructing a
legal execution trace, never by sampling cells independently, so every enforcement of the
reference holds on them."""
import numpy as np

from . import abi

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(seed: int, n: int, stream: int = 0) -> np.ndarray:
    """n outputs of SplitMix64 started at `seed` (vectorised; counter-based so streams are independent)"""
    with np.errstate(over="ignore"):
        i = np.arange(1, n + 1, dtype=np.uint64)
        z = np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + np.uint64((stream * 0xD1342543DE82EF95) & 0xFFFFFFFFFFFFFFFF)
        z = z + i * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def ram_trace(n: int, seed: int = 0xC1, n_cells: int = 1 << 10, n_nondet: int = 0,
              heap_page: int = 10):
    """C1: n MemoryQuery records over n_cells cells; returns (unsorted, sorted) record arrays.
    timestamps strictly increasing, first access of a cell is a write, later accesses 50 % reads
    (value = last write) / 50 % writes (256-bit uniform), is_ptr on 1 % of the writes;
    sorted order = (page, index, timestamp), the order ram_permutation/mod.rs:296-316 enforces.
    n_nondet extra bootloader-heap writes with timestamp 0 lead the unsorted queue (:260-290)."""
    m = n - n_nondet
    assert m >= 0
    q = np.zeros(n, dtype=abi.MEMORY_QUERY_DTYPE)
    r = splitmix64(seed, m, 0)
    n_cells = max(1, min(n_cells, max(m, 1)))
    cell = (r % np.uint64(n_cells)).astype(np.int64)
    ts = np.arange(1, m + 1, dtype=np.uint32)
    order = np.lexsort((ts, cell))  # by cell then time
    sc = cell[order]
    first = np.ones(m, dtype=bool)
    first[1:] = sc[1:] != sc[:-1]
    coin = (splitmix64(seed, m, 1) >> np.uint64(63)).astype(bool)
    is_write_sorted = first | coin[order]
    ptr_coin = (splitmix64(seed, m, 2) % np.uint64(100)) == 0
    vals = np.stack([splitmix64(seed, m, 3 + k) for k in range(4)], axis=1).view("<u4").reshape(m, 8)
    pos = np.where(is_write_sorted, np.arange(m), 0)
    last_write = np.maximum.accumulate(pos)  # first element of every cell group is a write
    src = order[last_write]  # original index of the governing write
    body = q[n_nondet:]
    body["timestamp"] = ts
    body["memory_page"] = 8 + (cell // 64) % 16 + 16 * (cell // 1024)
    body["index"] = cell % 64
    is_write = np.zeros(m, dtype=bool)
    is_write[order] = is_write_sorted
    body["rw_flag"] = is_write
    value = np.empty((m, 8), dtype=np.uint32)
    value[order] = vals[src]
    body["value"] = value
    is_ptr = np.zeros(m, dtype=bool)
    is_ptr[order] = ptr_coin[src]
    body["is_ptr"] = is_ptr
    if n_nondet:
        head = q[:n_nondet]
        head["timestamp"] = 0
        head["memory_page"] = heap_page
        head["index"] = 100000 + np.arange(n_nondet, dtype=np.uint32)
        head["rw_flag"] = 1
        head["value"] = splitmix64(seed, n_nondet * 4, 9).view("<u4").reshape(n_nondet, 8)
    srt = np.lexsort((q["timestamp"], q["index"], q["memory_page"]))
    return q, q[srt]


def events_trace(n: int, seed: int = 0xC4, rollback_pct: int = 10):
    """C4 (log_sorter): n event LogQuery records with unique timestamps, `rollback_pct` % of the forward
    events followed LATER in the unsorted queue by their rollback twin (same timestamp, rollback = 1);
    sorted order = by (timestamp, rollback), the order log_sorter/mod.rs:315-350 enforces.
    Every record is a write (rw_flag = 1, :295-297).  Returns (unsorted, sorted)."""
    q = np.zeros(n, dtype=abi.LOG_QUERY_DTYPE)
    if n == 0:
        return q, q.copy()
    want_rb = (splitmix64(seed, n, 0) % np.uint64(100)) < np.uint64(rollback_pct)
    # forward events take slots until n is reached together with their twins
    n_fwd = n
    while True:
        n_rb = int(want_rb[:n_fwd].sum())
        if n_fwd + n_rb <= n:
            break
        n_fwd -= max(1, (n_fwd + n_rb - n) // 2)
    n_rb = n - n_fwd
    rb_src = np.flatnonzero(want_rb[:n_fwd])[:n_rb]
    if len(rb_src) < n_rb:  # top up so the total is exactly n
        extra = np.setdiff1d(np.arange(n_fwd), rb_src)[:n_rb - len(rb_src)]
        rb_src = np.sort(np.concatenate([rb_src, extra]))
    fwd = q[:n_fwd]
    fwd["timestamp"] = 1000 + 4 * np.arange(n_fwd, dtype=np.uint32)
    fwd["address"] = splitmix64(seed, n_fwd * 3, 1).view("<u4").reshape(n_fwd, 6)[:, :5]
    fwd["key"] = splitmix64(seed, n_fwd * 4, 2).view("<u4").reshape(n_fwd, 8)
    fwd["written_value"] = splitmix64(seed, n_fwd * 4, 3).view("<u4").reshape(n_fwd, 8)
    r = splitmix64(seed, n_fwd, 4)
    fwd["tx_number_in_block"] = (r % np.uint64(1000)).astype(np.uint32)
    service = ((r >> np.uint64(20)) & np.uint64(1)).astype(np.uint32)
    aux = ((r >> np.uint64(24)) % np.uint64(3)).astype(np.uint32)
    fwd["flags"] = aux | (1 << 16) | (service << 18)
    q[n_fwd:] = fwd[rb_src]
    q["flags"][n_fwd:] |= 1 << 17
    # the unsorted queue is in execution order: rollbacks appear after all forward events (reverted frames)
    rollback = (q["flags"] >> 17) & 1
    srt = np.lexsort((rollback, q["timestamp"]))
    return q, q[srt]


def decommit_requests_trace(n: int, seed: int = 0xC4, n_hashes: int = 1 << 10):
    """(f)1 (sort_decommittment_requests): n DecommitQuery records over n_hashes distinct code hashes in execution
    order: unique ascending timestamps, the first request of a hash carries is_first = 1 and fixes the page every later
    request of that hash repeats (what add_to_decommittment_queue of the VM produces).  sorted order = by
    (code_hash as a 256-bit integer, timestamp): concatenate_key, sort_decommittment_requests/mod.rs:383-401.
    Returns (unsorted, sorted)."""
    q = np.zeros(n, dtype=abi.DECOMMIT_QUERY_DTYPE)
    if n == 0:
        return q, q.copy()
    n_hashes = max(1, min(n_hashes, n))
    hashes = splitmix64(seed, n_hashes * 4, 1).view("<u4").reshape(n_hashes, 8).copy()
    hashes[:, 7] |= 1  # never the all-zero placeholder hash
    which = (splitmix64(seed, n, 2) % np.uint64(n_hashes)).astype(np.int64)
    q["code_hash"] = hashes[which]
    q["timestamp"] = 1000 + 4 * np.arange(n, dtype=np.uint32)
    first_pos = np.full(n_hashes, n, dtype=np.int64)
    np.minimum.at(first_pos, which, np.arange(n))
    q["is_first"] = (first_pos[which] == np.arange(n)).astype(np.uint32)
    q["page"] = (2048 + 8 * first_pos[which]).astype(np.uint32)
    keys = [q["timestamp"]] + [q["code_hash"][:, i] for i in range(8)]
    return q, q[np.lexsort(keys)]


def vm_log_queue_trace(n: int, seed: int = 0xC4, mix=(50, 25, 10, 6, 6, 3)):
    """(f)2 (demux_log_queue): n LogQuery records as the VM's log opcode emits them, in execution order, `mix` = per cent
    of rollup-storage accesses, events, L2->L1 messages, keccak256 / sha256 / ecrecover precompile calls (aux byte and
    formal address as demux_log_queue/mod.rs:285-330 tests them).  Returns the records."""
    q = np.zeros(n, dtype=abi.LOG_QUERY_DTYPE)
    if n == 0:
        return q
    r = splitmix64(seed, n, 0) % np.uint64(100)
    edges = np.cumsum(mix)
    kind = np.searchsorted(edges, r.astype(np.int64), side="right").clip(0, 5)
    q["address"] = splitmix64(seed, n * 3, 1).view("<u4").reshape(n, 6)[:, :5]
    q["key"] = splitmix64(seed, n * 4, 2).view("<u4").reshape(n, 8)
    q["read_value"] = splitmix64(seed, n * 4, 3).view("<u4").reshape(n, 8)
    q["written_value"] = splitmix64(seed, n * 4, 4).view("<u4").reshape(n, 8)
    q["timestamp"] = 1000 + 4 * np.arange(n, dtype=np.uint32)
    t = splitmix64(seed, n, 5)
    q["tx_number_in_block"] = (t % np.uint64(1000)).astype(np.uint32)
    rw = ((t >> np.uint64(20)) & np.uint64(1)).astype(np.uint32)
    aux = np.array([abi.STORAGE_AUX_BYTE, abi.EVENT_AUX_BYTE, abi.L1_MESSAGE_AUX_BYTE, abi.PRECOMPILE_AUX_BYTE,
                    abi.PRECOMPILE_AUX_BYTE, abi.PRECOMPILE_AUX_BYTE], dtype=np.uint32)[kind]
    q["flags"] = aux | (np.where(kind == 0, rw, 1).astype(np.uint32) << 16)
    for k, addr in ((3, abi.KECCAK256_PRECOMPILE_ADDRESS), (4, abi.SHA256_PRECOMPILE_ADDRESS), (5, abi.ECRECOVER_PRECOMPILE_ADDRESS)):
        sel = kind == k
        q["address"][sel] = 0
        q["address"][sel, 0] = addr
    return q


def code_decommit_requests(n: int, seed: int = 0xC4, max_words: int = 63):
    """(f)3 (code_unpacker_sha256): n deduplicated decommitment requests (what sort_decommittment_requests leaves) with their
    bytecodes: an odd number of 32-byte words each (the padding block then completes the last SHA-256 round,
    code_unpacker_sha256/mod.rs:215-221), code_hash = SHA-256 of the code with the top 4 bytes replaced by the version byte
    and the length in words (versioned hash, mod.rs:187-213).  Returns (requests, code_words [total_words, 8] uint32 limbs)."""
    import hashlib
    q = np.zeros(n, dtype=abi.DECOMMIT_QUERY_DTYPE)
    words = []
    lens = 1 + 2 * (splitmix64(seed, max(n, 1), 0) % np.uint64((max_words + 1) // 2)).astype(np.int64)
    for i in range(n):
        code = splitmix64(seed + 1 + i, int(lens[i]) * 4, 1).tobytes()
        digest = int.from_bytes(hashlib.sha256(code).digest(), "big")
        q["code_hash"][i] = [(digest >> (32 * k)) & 0xFFFFFFFF for k in range(7)] + [(abi.CODE_HASH_VERSION_TOP16 << 16) | int(lens[i])]
        for wi in range(int(lens[i])):
            v = int.from_bytes(code[32 * wi:32 * wi + 32], "big")
            words.append([(v >> (32 * k)) & 0xFFFFFFFF for k in range(8)])
    q["page"] = 2048 + 8 * np.arange(n, dtype=np.uint32)
    q["timestamp"] = 1000 + 4 * np.arange(n, dtype=np.uint32)
    q["is_first"] = 1
    return q, np.array(words, dtype=np.uint32).reshape(-1, 8)


def storage_trace(n: int, seed: int = 0xC4, n_cells: int = 1 << 16, shard: int = 0, first_position: int = 0):
    """C4 (storage_validity): n storage LogQuery records over n_cells (address, key) cells: 60 % reads,
    30 % writes, 10 % write + rollback pairs (the rollback twin directly follows its write), shard 0.
    Per cell the history is consistent: a read returns the current value, a write records the value it
    replaces in read_value, a rollback restores it.  Returns (unsorted, sorted, sorted_timestamps) where
    sorted is ordered by (address, key) as a 13-limb little-endian integer and then by queue position, and
    sorted_timestamps = position of the record in the unsorted queue (the `cycle_idx` the circuit packs
    into the unsorted encoding, storage_validity_by_grand_product/mod.rs:585-610)."""
    q = np.zeros(n, dtype=abi.LOG_QUERY_DTYPE)
    if n == 0:
        return q, q.copy(), np.zeros(0, dtype=np.uint32)
    n_cells = max(1, min(n_cells, n))
    kind_r = splitmix64(seed, n, 0) % np.uint64(100)
    # build the op list: 0 = read, 1 = write, 2 = write of a pair, 3 = rollback of the previous record
    kind = np.where(kind_r < 60, 0, np.where(kind_r < 90, 1, 2)).astype(np.int8)
    is_pair_w = kind == 2
    is_pair_w[-1] = False
    # a pair occupies positions i and i + 1; drop overlapping pair starts
    nxt = np.zeros(n, dtype=bool); nxt[1:] = is_pair_w[:-1]
    is_pair_w &= ~nxt
    nxt = np.zeros(n, dtype=bool); nxt[1:] = is_pair_w[:-1]
    kind = np.where(is_pair_w, 2, np.where(nxt, 3, np.where(kind == 2, 1, kind))).astype(np.int8)
    cell = (splitmix64(seed, n, 1) % np.uint64(n_cells)).astype(np.int64)
    cell[kind == 3] = cell[np.flatnonzero(kind == 3) - 1]
    caddr = splitmix64(seed, n_cells * 3, 2).view("<u4").reshape(n_cells, 6)[:, :5].copy()
    ckey = splitmix64(seed, n_cells * 4, 3).view("<u4").reshape(n_cells, 8).copy()
    caddr[:, 0] |= 1  # never the all-zero key
    cinit = splitmix64(seed, n_cells * 4, 4).view("<u4").reshape(n_cells, 8)
    newval = splitmix64(seed, n * 4, 5).view("<u4").reshape(n, 8)
    pos = np.arange(n)
    order = np.lexsort((pos, cell))
    sc, sk = cell[order], kind[order]
    first = np.ones(n, dtype=bool); first[1:] = sc[1:] != sc[:-1]
    # value before each record = written value of the last PLAIN write strictly before it in the cell, else the
    # cell's initial value
    is_plain_w = sk == 1
    idx = np.where(is_plain_w, np.arange(n), -1)
    seg_start = np.maximum.accumulate(np.where(first, np.arange(n), 0))
    last_w_incl = np.maximum.accumulate(idx)
    last_w_excl = np.empty(n, dtype=np.int64); last_w_excl[0] = -1; last_w_excl[1:] = last_w_incl[:-1]
    has_w = last_w_excl >= seg_start
    before = np.where(has_w[:, None], newval[order][np.maximum(last_w_excl, 0)], cinit[sc])
    written = np.where((sk == 0)[:, None], before, newval[order])
    rb_rows = np.flatnonzero(sk == 3)
    written[rb_rows] = written[rb_rows - 1]
    rec = np.zeros(n, dtype=abi.LOG_QUERY_DTYPE)
    rec["address"], rec["key"] = caddr[sc], ckey[sc]
    rec["read_value"], rec["written_value"] = before, written
    rec["timestamp"] = 100 + order
    rec["flags"] = (shard << 8) | ((sk != 0).astype(np.uint32) << 16) | ((sk == 3).astype(np.uint32) << 17)
    q[order] = rec
    # sorted: by the 13-limb packed key (address most significant) then by queue position
    crank_order = np.lexsort(tuple(ckey[:, i] for i in range(8)) + tuple(caddr[:, i] for i in range(5)))
    crank = np.empty(n_cells, dtype=np.int64); crank[crank_order] = np.arange(n_cells)
    srt = np.lexsort((pos, crank[cell]))
    return q, q[srt], (first_position + srt).astype(np.uint32)


def bytes_to_u256_words(data: bytes, unalignment: int) -> np.ndarray:
    """memory words of a byte string that starts `unalignment` bytes into its first 32-byte word (big-endian words,
    0xff filler before the data, zero fill after it) -- the reference test's helper,
    keccak256_round_function/mod.rs:971-998.  Returns [n_words, 8] little-endian u32 limbs."""
    stream = b"\xff" * unalignment + data
    if len(stream) % 32:
        stream += bytes(32 - len(stream) % 32)
    words = np.frombuffer(stream, dtype=">u4").reshape(-1, 8)[:, ::-1]
    return np.ascontiguousarray(words.astype("<u4"))


def precompile_call(address: int, in_offset: int, in_length: int, out_offset: int, in_page: int, out_page: int,
                    timestamp: int, extra: int = 0, aux_byte: int = abi.PRECOMPILE_AUX_BYTE):
    """LogQuery of a precompile call; key = PrecompileCallABI::to_u256 (zkevm_opcode_defs, un-vendored): limbs
    [0] input offset, [1] input length, [2] output offset, [3] output length, [4] page to read, [5] page to write,
    [6..8] precompile_interpreted_data (keccak256_round_function/mod.rs:74-81, sha256_round_function/mod.rs:66-72)"""
    q = np.zeros((), dtype=abi.LOG_QUERY_DTYPE)
    q["address"][0] = address
    q["key"] = [in_offset, in_length, out_offset, 1, in_page, out_page, extra & 0xFFFFFFFF, extra >> 32]
    q["timestamp"] = timestamp
    q["flags"] = abi.lq_flags(aux=aux_byte, rw=1)
    return q


def keccak_calls(n_calls: int, seed: int = 0xC3, max_len: int = 1024):
    """C3 (keccak): back-to-back precompile calls with input lengths uniform in [0, max_len) bytes and unalignment
    uniform in [0, 32).  Returns (requests [n_calls], memory_reads [n_words, 8], messages: list of bytes)."""
    r = splitmix64(seed, 2 * n_calls, 0)
    lengths = (r[:n_calls] % np.uint64(max(1, max_len))).astype(np.int64)
    unal = (r[n_calls:] % np.uint64(32)).astype(np.int64)
    blob = splitmix64(seed, int(lengths.sum()) // 8 + 1, 1).tobytes()
    reqs = np.zeros(n_calls, dtype=abi.LOG_QUERY_DTYPE)
    words, msgs, off = [], [], 0
    for i in range(n_calls):
        msg = blob[off:off + int(lengths[i])]
        off += int(lengths[i])
        msgs.append(msg)
        base_word = 1000 * i
        reqs[i] = precompile_call(abi.KECCAK256_PRECOMPILE_ADDRESS, base_word * 32 + int(unal[i]), len(msg), 7 + i,
                                  100 + i, 200 + i, 10 + 4 * i)
        if len(msg):
            words.append(bytes_to_u256_words(msg, int(unal[i])))
    reads = np.concatenate(words) if words else np.zeros((0, 8), dtype=np.uint32)
    return reqs, reads, msgs


def sha256_pad(msg: bytes) -> bytes:
    """FIPS 180-4 padding: the sha256 circuit consumes pre-padded blocks (sha256_round_function/mod.rs:72)"""
    ml = len(msg) * 8
    msg = msg + b"\x80"
    msg += bytes((56 - len(msg) % 64) % 64)
    return msg + ml.to_bytes(8, "big")


def sha256_calls(n_calls: int, seed: int = 0xC3, max_rounds: int = 16):
    """C3 (sha256): back-to-back precompile calls with num_rounds uniform in [1, max_rounds] (64-byte blocks of an
    already padded message).  Returns (requests, memory_reads [2 * total_rounds, 8], messages)."""
    r = splitmix64(seed, n_calls, 0)
    rounds = (1 + r % np.uint64(max_rounds)).astype(np.int64)
    reqs = np.zeros(n_calls, dtype=abi.LOG_QUERY_DTYPE)
    words, msgs = [], []
    rr = splitmix64(seed, n_calls, 1)
    blob = splitmix64(seed, int(rounds.sum()) * 8 + 8, 2).tobytes()
    off = 0
    for i in range(n_calls):
        # a message whose padded length is exactly rounds[i] blocks
        max_len = int(rounds[i]) * 64 - 9
        min_len = max(0, (int(rounds[i]) - 1) * 64 - 8)
        ln = min_len + int(rr[i] % np.uint64(max_len - min_len + 1))
        msg = blob[off:off + ln]
        off += ln
        padded = sha256_pad(msg)
        assert len(padded) == int(rounds[i]) * 64
        msgs.append(msg)
        reqs[i] = precompile_call(abi.SHA256_PRECOMPILE_ADDRESS, 50 * i, 0, 9 + i, 300 + i, 400 + i, 20 + 4 * i,
                                  extra=int(rounds[i]))
        words.append(bytes_to_u256_words(padded, 0))
    return reqs, np.concatenate(words), msgs
