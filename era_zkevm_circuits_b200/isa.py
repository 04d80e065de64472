"""Synthetic EraVM ISA tables, assembler and program generator for main_vm (host-side input preparation).

`zkevm_opcode_defs` (the crate that owns the real ISA tables) is un-vendored, so the engine takes the tables as input
data (`abi.VmIsa`): opcode -> (price, 48-bit property bit spread + 3 aux bits), the layout of
/root/reference/src/tables/opcodes_decoding.rs:14-38 and main_vm/opcode_bitmask.rs:83-127.  This module builds a
table with that exact layout whose opcode NUMBERING is synthetic (sequential), plus the 64-bit opcode word encoding of
main_vm/decoded_opcode.rs:408-514: bits 0..11 variant, 13..16 condition, 16..24 src registers, 24..32 dst registers,
32..48 imm0, 48..64 imm1."""
import numpy as np

from . import abi
from .synthetic import splitmix64

(OP_INVALID, OP_NOP, OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_JUMP, OP_CONTEXT, OP_SHIFT, OP_BINOP, OP_PTR, OP_NEAR_CALL, OP_LOG,
 OP_FAR_CALL, OP_RET, OP_UMA) = range(16)
MODE_REG, MODE_PUSH_POP, MODE_STACK_OFFSET, MODE_STACK_ABS, MODE_IMM16, MODE_CODE = range(6)
COND_ALWAYS, COND_GT, COND_LT, COND_EQ, COND_GE, COND_LE, COND_NE, COND_GT_OR_LT = range(8)
TYPE_BITS, VARIANT_BITS, FLAG_BITS, SRC_BITS, DST_BITS = 16, 10, 2, 6, 4
FULL_DST = (OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_SHIFT, OP_BINOP, OP_PTR)  # can_write_dst0_into_memory
N_VARIANTS = {OP_CONTEXT: 10, OP_SHIFT: 4, OP_BINOP: 3, OP_PTR: 4}
LOG_STORAGE_READ, LOG_STORAGE_WRITE, LOG_TO_L1, LOG_EVENT, LOG_PRECOMPILE = range(5)
RET_OK, RET_REVERT, RET_PANIC = range(3)
UMA_HEAP_READ, UMA_HEAP_WRITE, UMA_AUX_READ, UMA_AUX_WRITE, UMA_PTR_READ = range(5)
FORWARD_HEAP, FORWARD_PTR, FORWARD_AUX = 0, 1, 2
FAR_CALL_NORMAL, FAR_CALL_DELEGATE, FAR_CALL_MIMIC = range(3)


def props_bits(op, variant=0, flags=0, src=MODE_REG, dst=MODE_REG):
    v = 1 << op
    v |= 1 << (TYPE_BITS + variant)
    v |= flags << (TYPE_BITS + VARIANT_BITS)
    v |= 1 << (TYPE_BITS + VARIANT_BITS + FLAG_BITS + src)
    v |= 1 << (TYPE_BITS + VARIANT_BITS + FLAG_BITS + SRC_BITS + dst)
    return v


class Isa:
    """table + reverse map (op, variant, flags, src, dst) -> opcode number"""

    def __init__(self):
        self.isa = abi.VmIsa()
        self.index = {}
        nxt = [0]

        def add(op, variant, flags, src, dst, price, kernel=0, static_ok=1, panic=0):
            i = nxt[0]
            nxt[0] += 1
            self.isa.opcode_price[i] = price
            self.isa.opcode_props[i] = props_bits(op, variant, flags, src, dst) | (kernel << 48) | (static_ok << 49) | (panic << 50)
            self.index[(op, variant, flags, src, dst)] = i
            return i

        add(OP_INVALID, 0, 0, MODE_REG, MODE_REG, 0, panic=1)
        nop = add(OP_NOP, 0, 0, MODE_REG, MODE_REG, 1)
        for op in (OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_SHIFT, OP_BINOP, OP_PTR):
            for variant in range(N_VARIANTS.get(op, 1)):
                for flags in range(4):
                    for src in range(6):
                        for dst in range(4):
                            add(op, variant, flags, src, dst, {OP_MUL: 3, OP_DIV: 4}.get(op, 2))
        for src in range(6):
            add(OP_JUMP, 0, 0, src, MODE_REG, 2)
        for variant in range(10):
            add(OP_CONTEXT, variant, 0, MODE_REG, MODE_REG, 2, kernel=int(variant >= 7), static_ok=int(variant < 7))
        for variant in (RET_OK, RET_REVERT, RET_PANIC):
            for flags in (0, 1):  # to-label
                add(OP_RET, variant, flags, MODE_REG, MODE_REG, 5)
        ret_panic = self.index[(OP_RET, RET_PANIC, 0, MODE_REG, MODE_REG)]
        add(OP_NEAR_CALL, 0, 0, MODE_REG, MODE_REG, 25)
        for variant in range(5):
            for flags in (0, 1):  # increment (reads) / increment (writes)
                add(OP_UMA, variant, flags, MODE_REG, MODE_REG, 6, static_ok=int(variant in (UMA_HEAP_READ, UMA_AUX_READ, UMA_PTR_READ)))
        for variant in range(5):
            for flags in (0, 1):  # first message
                add(OP_LOG, variant, flags, MODE_REG, MODE_REG, 40, kernel=int(variant in (LOG_TO_L1, LOG_PRECOMPILE)),
                    static_ok=int(variant == LOG_STORAGE_READ))
        for variant in range(3):      # normal / delegate / mimic
            for flags in range(4):    # static | shard << 1
                add(OP_FAR_CALL, variant, flags, MODE_REG, MODE_REG, 100, static_ok=1)
        assert nxt[0] <= 2048
        for i in range(nxt[0], 2048):  # unused opcode numbers decode to Invalid with the explicit-panic aux bit
            self.isa.opcode_props[i] = self.isa.opcode_props[0]
        for c in range(8):
            for f in range(8):
                of, eq, gt = f & 1, (f >> 1) & 1, (f >> 2) & 1
                self.isa.condition_table[c][f] = int([True, gt, of, eq, gt or eq, of or eq, not eq, gt or of][c])
        self.isa.nop_opcode_encoding = nop
        self.isa.panic_opcode_encoding = ret_panic
        self.isa.nop_bitspread = self.isa.opcode_props[nop] & ((1 << 48) - 1)
        self.isa.panic_bitspread = self.isa.opcode_props[ret_panic] & ((1 << 48) - 1)
        # zkevm_opcode_defs::system_params (from memory; data, not logic)
        self.isa.bootloader_base_page, self.isa.bootloader_code_page, self.isa.bootloader_calldata_page = 8, 8, 7
        self.isa.starting_timestamp, self.isa.starting_base_page = 1024, 16  # far calls take pages from here (bootloader: 8..11)
        self.isa.initial_frame_formal_eh_location, self.isa.vm_initial_frame_ergs = 0xFFFF, 0xFFFFFFFF
        self.isa.bootloader_formal_address_low, self.isa.bootloader_max_memory = 0x8001, 1 << 24
        self.isa.vm_max_stack_depth = 1 << 16
        for i, v in enumerate((0, 1, 2, 3)):  # STORAGE / EVENT / L1_MESSAGE / PRECOMPILE aux bytes
            self.isa.log_aux_bytes[i] = v
        self.isa.initial_storage_write_pubdata_bytes, self.isa.l1_message_pubdata_bytes = 64, 88
        # far call parameters (zkevm_opcode_defs, from memory; data, not logic)
        self.isa.new_frame_memory_stipend, self.isa.new_memory_pages_per_far_call = 1 << 12, 8
        self.isa.deployer_system_contract_address_low, self.isa.ergs_per_code_word_decommittment = 0x8006, 4
        self.isa.code_hash_version_byte, self.isa.code_hash_yet_constructed_marker, self.isa.code_hash_at_rest_marker = 1, 1, 0
        self.isa.call_system_abi_registers[0], self.isa.call_system_abi_registers[1] = 2, 12   # r3..r12
        self.isa.call_reserved_range[0], self.isa.call_reserved_range[1] = 12, 14              # r13, r14
        self.isa.call_implicit_parameter_reg_idx = 14                                         # r15

    def encode(self, op, variant=0, flags=0, src=MODE_REG, dst=MODE_REG, cond=COND_ALWAYS, src0=0, src1=0, dst0=0, dst1=0,
               imm0=0, imm1=0):
        num = self.index[(op, variant, flags, src, dst)]
        return num | (cond << 13) | (src0 << 16) | (src1 << 20) | (dst0 << 24) | (dst1 << 28) | (imm0 << 32) | (imm1 << 48)


def pack_code(opcodes):
    """4 opcodes per 32-byte word, big-endian: sub-pc 0 -> limbs [6],[7] ... sub-pc 3 -> limbs [0],[1]
    (main_vm/pre_state.rs:185-206).  Returns [n_words, 8] uint32."""
    ops = list(opcodes)
    while len(ops) % 4:
        ops.append(0)
    words = np.zeros((len(ops) // 4, 8), dtype=np.uint32)
    for i, o in enumerate(ops):
        w, sub = divmod(i, 4)
        words[w, 6 - 2 * sub] = o & 0xFFFFFFFF
        words[w, 7 - 2 * sub] = o >> 32
    return words


def random_program(isa: Isa, n: int, seed: int = 0xC2, with_memory=True, full=True, far_calls=False):
    """the C2 instruction mix of SURVEY 8d: 40 % add/sub, 15 % binop, 10 % mul/div, 10 % shifts, 10 % UMA heap / aux heap
    reads and writes, 5 % jumps, 5 % context / ptr, 3 % log (storage reads / writes, events, L1 messages, precompile calls),
    2 % near calls into small subroutines that return ok / revert / panic; 30 % of the arithmetic uses stack / code-page /
    immediate operands, 10 % of the instructions are conditional.  `n` instructions: a main body that ends with a jump
    back to 0 (so the program runs for any number of cycles) followed by the subroutines.  r14 / r15 are reserved for
    storage keys and heap offsets (each UMA / log is preceded by the immediate load of its address, counted as an add).
    full=False restricts the mix to the arithmetic / addressing subset (no uma / log / calls).  far_calls=True (not part
    of the C2 mix) turns a quarter of the calls into far calls of a deployed address: every frame runs this same program
    from pc 0, so it opens with a 3-instruction dispatch (caller == 0: the root's main body; otherwise the callee body,
    which logs, touches its own heap and returns)."""
    r = splitmix64(seed, 8 * n, 0).reshape(n, 8)
    n_subs = max(1, n // 64) if full else 0
    sub_len = 6
    callee_len = 8 if far_calls else 0
    n_main = n - n_subs * sub_len - callee_len
    assert n_main >= (16 if far_calls else 8)
    ops = []
    if far_calls:
        ops += [isa.encode(OP_CONTEXT, 1, 0, dst0=13),                                   # r13 = caller
                isa.encode(OP_SUB, 0, 1, src0=13, src1=0, dst0=13),                      # EQ iff the root frame
                isa.encode(OP_JUMP, 0, 0, MODE_IMM16, cond=COND_NE, imm0=n_main)]        # a callee: its body sits behind the main body

    def arith(i, k, allow_mem=True):
        src0, src1 = (int(r[i, j] % 14) + 2 for j in (1, 2))
        dst0, dst1 = (int(r[i, j] % 12) + 2 for j in (3, 4))  # r2..r13: r1 keeps the calldata pointer, r14 / r15 are reserved
        flags = int(r[i, 5] % 4)
        cond = COND_ALWAYS if int(r[i, 6] % 10) else int(r[i, 6] >> 8) % 8
        src, dst, imm0, imm1 = MODE_REG, MODE_REG, 0, 0
        if with_memory and allow_mem and int(r[i, 7] % 10) < 3:
            m = int(r[i, 7] >> 8) % 5
            if m == 0:
                src, imm0 = MODE_IMM16, int(r[i, 7] >> 16) & 0xFFFF
            elif m == 1:
                src, src0, imm0 = MODE_CODE, 0, int(r[i, 7] >> 16) % max(1, n // 4)
            elif m == 2:
                src, src0, imm0 = MODE_STACK_ABS, 0, int(r[i, 7] >> 16) % 64
            elif m == 3:
                dst, dst0, imm1 = MODE_STACK_ABS, 0, int(r[i, 7] >> 16) % 64
            else:
                src, src0, imm0, dst, dst0, imm1 = MODE_STACK_OFFSET, 0, int(r[i, 7] >> 16) % 16, MODE_PUSH_POP, 0, 1
        if k < 40:
            return isa.encode(OP_ADD if k % 2 else OP_SUB, 0, flags, src, dst, cond, src0, src1, dst0, imm0=imm0, imm1=imm1)
        if k < 55:
            return isa.encode(OP_BINOP, k % 3, flags, src, dst, cond, src0, src1, dst0, imm0=imm0, imm1=imm1)
        if k < 65:
            return isa.encode(OP_MUL if k % 2 else OP_DIV, 0, flags, src, dst, cond, src0, src1, dst0, dst1, imm0, imm1)
        return isa.encode(OP_SHIFT, k % 4, flags, src, dst, cond, src0, src1, dst0, imm0=imm0, imm1=imm1)

    i = 0
    while len(ops) < n_main - 1:
        room = n_main - 1 - len(ops)
        k = int(r[i, 0] % 100) if full else int(r[i, 0] % 75) if int(r[i, 0] % 100) < 92 else 80 + int(r[i, 0] % 6)
        x = int(r[i, 7])
        dst0 = int(r[i, 3] % 12) + 2
        if k < 75:
            ops.append(arith(i, k))
        elif k < 80:  # short forward jump
            ops.append(isa.encode(OP_JUMP, 0, 0, MODE_IMM16, imm0=min(len(ops) + 1 + x % 3, n_main - 1)))
        elif k < 83:  # ptr.add of the calldata pointer by a small immediate never panics
            ops.append(isa.encode(OP_PTR, 0, 1, MODE_IMM16, MODE_REG, COND_ALWAYS, src1=1, dst0=dst0, imm0=x % 7))
        elif k < 85:
            # set_ergs_per_pubdata takes the (small) key register: a random 256-bit price would burn every erg at the next write
            variant = x % 10
            ops.append(isa.encode(OP_CONTEXT, variant, 0, MODE_REG, MODE_REG, COND_ALWAYS, src0=14 if variant == 8 else int(r[i, 1] % 14) + 2, dst0=dst0))
        elif k < 95 and room >= 2:  # UMA: offset -> r15, then the access
            variant = (UMA_HEAP_READ, UMA_HEAP_WRITE, UMA_HEAP_READ, UMA_HEAP_WRITE, UMA_AUX_READ, UMA_AUX_WRITE, UMA_PTR_READ)[x % 7]
            inc = (x >> 3) & 1
            off = (x >> 8) % 4096 if (x >> 4) & 1 else ((x >> 8) % 128) * 32  # half of the accesses are aligned
            if variant == UMA_PTR_READ:  # through the (empty) calldata pointer: a legitimate read beyond the slice
                ops.append(isa.encode(OP_UMA, variant, inc, src0=1, dst0=dst0, dst1=13))
            else:
                ops.append(isa.encode(OP_ADD, 0, 0, MODE_IMM16, src1=0, dst0=15, imm0=off))
                ops.append(isa.encode(OP_UMA, variant, inc, src0=15, src1=int(r[i, 2] % 14) + 2, dst0=dst0, dst1=15))
        elif k < 98 and room >= 2:  # log: key -> r14, then the query
            variant = (LOG_STORAGE_READ, LOG_STORAGE_WRITE, LOG_STORAGE_WRITE, LOG_EVENT, LOG_TO_L1, LOG_PRECOMPILE)[x % 6]
            ops.append(isa.encode(OP_ADD, 0, 0, MODE_IMM16, src1=0, dst0=14, imm0=(x >> 8) % 48))
            src1 = 14 if variant == LOG_PRECOMPILE else int(r[i, 2] % 14) + 2  # a precompile call burns src1[0] ergs
            ops.append(isa.encode(OP_LOG, variant, (x >> 3) & 1 if variant in (LOG_EVENT, LOG_TO_L1) else 0, src0=14, src1=src1, dst0=dst0))
        elif k >= 98 and far_calls and (x >> 20) % 4 == 0 and room >= 5:
            # ABI register: ergs in bits 192..224 (forwarding mode 0, normal call); every odd address is deployed
            ops.append(isa.encode(OP_ADD, 0, 0, MODE_IMM16, src1=0, dst0=13, imm0=20000 + x % 30000))
            ops.append(isa.encode(OP_ADD, 0, 0, MODE_IMM16, src1=0, dst0=15, imm0=192))
            ops.append(isa.encode(OP_SHIFT, 0, 0, src0=13, src1=15, dst0=13))
            ops.append(isa.encode(OP_ADD, 0, 0, MODE_IMM16, src1=0, dst0=14, imm0=0x9000 | ((x >> 8) % 64) | ((x >> 16) & 1)))
            ops.append(isa.encode(OP_FAR_CALL, (x >> 24) % 3, (x >> 28) % 4 & 1, src0=13, src1=14, imm0=len(ops) + 1))
        elif k >= 98 and n_subs:
            sub = n_main + callee_len + (x % n_subs) * sub_len
            ops.append(isa.encode(OP_NEAR_CALL, src0=0, imm0=sub, imm1=len(ops) + 1))
        else:
            ops.append(isa.encode(OP_NOP, cond=int(r[i, 6] >> 8) % 8))
        i += 1
    ops.append(isa.encode(OP_JUMP, 0, 0, MODE_IMM16, imm0=0))
    if far_calls:  # the callee body: registers were cleaned by the call, r0-based operands only
        ops += [isa.encode(OP_ADD, 0, 0, MODE_IMM16, src1=0, dst0=2, imm0=int(r[0, 1] % 1000)),
                isa.encode(OP_ADD, 0, 0, MODE_IMM16, src1=0, dst0=14, imm0=int(r[0, 2] % 48)),
                isa.encode(OP_LOG, LOG_STORAGE_WRITE, 0, src0=14, src1=2),
                isa.encode(OP_UMA, UMA_HEAP_WRITE, 1, src0=14, src1=2, dst0=15),
                isa.encode(OP_UMA, UMA_HEAP_READ, 0, src0=14, dst0=3),
                isa.encode(OP_MUL, 0, 0, src0=2, src1=3, dst0=4, dst1=5),
                isa.encode(OP_LOG, LOG_EVENT, 0, src0=14, src1=4),
                isa.encode(OP_RET, RET_OK if int(r[0, 3] % 4) else RET_REVERT)]
    for sidx in range(n_subs):
        j = n_main + callee_len + sidx * sub_len
        x = int(r[j, 7])
        body = [arith(j + t, int(r[j + t, 0] % 75), allow_mem=False) for t in range(sub_len - 3)]
        body.append(isa.encode(OP_ADD, 0, 0, MODE_IMM16, src1=0, dst0=14, imm0=(x >> 8) % 48))
        body.append(isa.encode(OP_LOG, (LOG_STORAGE_WRITE, LOG_EVENT)[x % 2], 0, src0=14, src1=int(r[j, 2] % 14) + 2))
        kind = x % 10
        body.append(isa.encode(OP_RET, RET_OK if kind < 6 else RET_REVERT if kind < 9 else RET_PANIC))
        ops += body
    assert len(ops) == n
    return ops
