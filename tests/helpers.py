"""Shared builders for the ram_permutation tests (inputs are built like the reference test does:
push everything into two empty queues, /root/reference/src/ram_permutation/mod.rs:506-515)."""
import ctypes as C

import numpy as np

import orc as O
from era_zkevm_circuits_b200 import abi


def ram_instance(orc, unsorted, sorted_, nondet_len=0):
    uprev, ufin = O.memory_queue_simulate(orc, unsorted)
    sprev, sfin = O.memory_queue_simulate(orc, sorted_)
    io = O.ram_closed_form(ufin, sfin, start=True, nondet_len=nondet_len)
    return io, uprev, sprev


def fsm_equal(a, b):
    return bytes(a) == bytes(b)


def continue_io(io_done):
    """closed-form input of the next chained instance: observable input carried over, hidden FSM input :=
    previous hidden FSM output (the reference's instance chaining, ram_permutation/input.rs:52-62)"""
    nxt = abi.RamClosedForm()
    nxt.start_flag = 0
    nxt.observable_input = io_done.observable_input
    C.memmove(C.byref(nxt.hidden_fsm_input), C.byref(io_done.hidden_fsm_output), C.sizeof(abi.RamFsm))
    return nxt
