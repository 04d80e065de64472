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
        ret_panic = add(OP_RET, 2, 0, MODE_REG, MODE_REG, 1)
        for op in (OP_NEAR_CALL, OP_LOG, OP_FAR_CALL, OP_UMA):
            add(op, 0, 0, MODE_REG, MODE_REG, 5, static_ok=0)
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
        self.isa.starting_timestamp, self.isa.starting_base_page = 1024, 8
        self.isa.initial_frame_formal_eh_location, self.isa.vm_initial_frame_ergs = 0xFFFF, 0xFFFFFFFF
        self.isa.bootloader_formal_address_low, self.isa.bootloader_max_memory = 0x8001, 1 << 24
        self.isa.vm_max_stack_depth = 1 << 16

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


def random_program(isa: Isa, n: int, seed: int = 0xC2, with_memory=True):
    """straight-line mix (restricted to the built opcode subset): 45 % add/sub, 17 % binop, 12 % mul/div,
    12 % shifts, 6 % ptr, 6 % context, 2 % nop/conditional; 30 % of the arithmetic uses stack / code-page / immediate
    operands.  The last instruction jumps back to 0, so the program runs for any number of cycles."""
    r = splitmix64(seed, 8 * n, 0).reshape(n, 8)
    ops = []
    for i in range(n - 1):
        k = int(r[i, 0] % 100)
        src0, src1, dst0, dst1 = (int(r[i, j] % 14) + 2 for j in (1, 2, 3, 4))  # r2..r15: r1 keeps the calldata pointer
        flags = int(r[i, 5] % 4)
        cond = COND_ALWAYS if int(r[i, 6] % 10) else int(r[i, 6] >> 8) % 8
        src, dst, imm0, imm1 = MODE_REG, MODE_REG, 0, 0
        if with_memory and int(r[i, 7] % 10) < 3:
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
        if k < 45:
            ops.append(isa.encode(OP_ADD if k % 2 else OP_SUB, 0, flags, src, dst, cond, src0, src1, dst0, imm0=imm0, imm1=imm1))
        elif k < 62:
            ops.append(isa.encode(OP_BINOP, k % 3, flags, src, dst, cond, src0, src1, dst0, imm0=imm0, imm1=imm1))
        elif k < 74:
            ops.append(isa.encode(OP_MUL if k % 2 else OP_DIV, 0, flags, src, dst, cond, src0, src1, dst0, dst1, imm0, imm1))
        elif k < 86:
            ops.append(isa.encode(OP_SHIFT, k % 4, flags, src, dst, cond, src0, src1, dst0, imm0=imm0, imm1=imm1))
        elif k < 92:  # ptr.add / ptr.shrink of the calldata pointer by a small immediate never panics
            ops.append(isa.encode(OP_PTR, 0, 1, MODE_IMM16, MODE_REG, COND_ALWAYS, src1=1, dst0=dst0, imm0=int(r[i, 7] % 7)))
        elif k < 98:
            ops.append(isa.encode(OP_CONTEXT, int(r[i, 7] % 10), 0, MODE_REG, MODE_REG, COND_ALWAYS, src0=src0, dst0=dst0))
        else:
            ops.append(isa.encode(OP_NOP, cond=cond))
    ops.append(isa.encode(OP_JUMP, 0, 0, MODE_IMM16, imm0=0))
    return ops
