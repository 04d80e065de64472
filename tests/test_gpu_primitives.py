"""CUDA primitives through the C ABI against the CPU oracle, bit-exact."""
import numpy as np
import pytest

import orc as O

P = 0xFFFFFFFF00000001
pytestmark = pytest.mark.gpu


def rand_field(rng, shape):
    return (rng.integers(0, 1 << 63, size=shape, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=shape, dtype=np.uint64)) % np.uint64(P)


def test_field_ops_edge_values_and_random(engine):
    edge = [0, 1, 2, P - 1, P - 2, 0xFFFFFFFF, 0x100000000, 0xFFFFFFFE, 0xFFFFFFFF00000000, 0xFFFFFFFEFFFFFFFF,
            0x8000000000000000, 0x7FFFFFFFFFFFFFFF, 0xFFFFFFFE00000001, 0x00000001FFFFFFFF, (1 << 63) % P]
    rng = np.random.default_rng(5)
    a = [x for x in edge for _ in edge] + rand_field(rng, (4000,)).tolist()
    b = [y for _ in edge for y in edge] + rand_field(rng, (4000,)).tolist()
    c = [edge[(i * 7) % len(edge)] for i in range(len(edge) ** 2)] + rand_field(rng, (4000,)).tolist()
    mul, add, sub, fma = engine.field_ops(np.array(a, dtype=np.uint64), np.array(b, dtype=np.uint64), np.array(c, dtype=np.uint64))
    for i, (x, y, z) in enumerate(zip(a, b, c)):
        x, y, z = int(x), int(y), int(z)
        assert int(mul[i]) == x * y % P, (hex(x), hex(y))
        assert int(add[i]) == (x + y) % P, (hex(x), hex(y))
        assert int(sub[i]) == (x - y) % P, (hex(x), hex(y))
        assert int(fma[i]) == (x * y + z) % P, (hex(x), hex(y), hex(z))


def test_poseidon2_batch(engine, orc):
    rng = np.random.default_rng(1)
    st = rand_field(rng, (1000, 12))
    st[0] = 0
    st[1] = P - 1
    got = engine.poseidon2_permute(st)
    assert np.array_equal(got, O.poseidon2(orc, st))


def test_poseidon2_device_buffers(engine, orc):
    import torch
    rng = np.random.default_rng(2)
    st = rand_field(rng, (4096 + 17, 12))
    d = torch.from_numpy(st.view(np.int64)).cuda()
    out = engine.poseidon2_permute(d)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy().view(np.uint64)[:64], O.poseidon2(orc, st[:64]))
    assert np.array_equal(out.cpu().numpy().view(np.uint64)[-3:], O.poseidon2(orc, st[-3:]))


@pytest.mark.parametrize("ln", [0, 1, 7, 8, 9, 18, 51, 69])
def test_commit_encoding(engine, orc, ln):
    rng = np.random.default_rng(ln)
    x = rand_field(rng, (5, ln))
    got = engine.commit_encoding(x)
    for i in range(5):
        assert got[i].tolist() == O.commit_encoding(orc, x[i]).tolist()


@pytest.mark.parametrize("enc_len,rows", [(8, 1), (8, 255), (8, 256), (8, 257), (8, 70000), (20, 1000), (20, 66000)])
def test_accumulate_grand_products(engine, orc, enc_len, rows):
    rng = np.random.default_rng(rows + enc_len)
    lhs, rhs = rand_field(rng, (enc_len, rows)), rand_field(rng, (enc_len, rows))
    ch = rand_field(rng, (2, enc_len + 1))
    ch[:, 0] = 1
    acc_in = rand_field(rng, (4,))
    flags = (rng.integers(0, 4, size=rows) != 0).astype(np.uint8)
    want = O.accumulate_grand_products(orc, lhs, rhs, ch, acc_in, flags, want_chain=True)
    got = engine.accumulate_grand_products(lhs, rhs, ch, acc_in, flags, want_chain=True)
    assert np.array_equal(got[2], want[2])
    assert np.array_equal(got[0], want[0])
    assert np.array_equal(got[1], want[1])
    # flags = None means "always accumulate"
    want = O.accumulate_grand_products(orc, lhs, rhs, ch, acc_in, None)
    got = engine.accumulate_grand_products(lhs, rhs, ch, acc_in, None)
    assert np.array_equal(got[2], want[2]) and np.array_equal(got[0], want[0])


def test_grand_product_permutation_invariance(engine):
    """size-independent property at the C4 size (2^22 rows): permuting the rows of the rhs side leaves the
    final products equal, and lhs == rhs when rhs is a permutation of lhs."""
    import torch
    rows, enc = 1 << 22, 20
    g = torch.Generator(device="cuda").manual_seed(3)
    lhs = torch.randint(0, 1 << 62, (enc, rows), generator=g, device="cuda", dtype=torch.int64)
    perm = torch.randperm(rows, generator=g, device="cuda")
    rhs = lhs[:, perm].contiguous()
    ch = rand_field(np.random.default_rng(4), (2, enc + 1))
    acc, _, fin = engine.accumulate_grand_products(lhs, rhs, ch, np.ones(4, dtype=np.uint64))
    assert fin[0] == fin[2] and fin[1] == fin[3] and fin[0] != fin[1]
    torch.cuda.synchronize()
    a = acc.cpu().numpy().view(np.uint64)
    assert a[:, -1].tolist() == fin.tolist()
    assert a.max() < P


def test_scale_accumulators_host_and_device(engine):
    from era_zkevm_circuits_b200 import abi
    """zkc_scale_accumulators: the 4-column fix-up multiply of a sharded grand product, against Python integers"""
    import torch
    P = abi.GL_P
    rng = np.random.default_rng(3)
    for rows in (1, 2, 7, 1000, 4097):
        acc = rng.integers(0, P, size=(4, rows), dtype=np.uint64)
        acc[0, 0] = P - 1
        f = np.array([P - 1, 1, 0x123456789ABCDEF % P, 2], dtype=np.uint64)
        want = np.array([[int(v) * int(f[c]) % P for v in acc[c]] for c in range(4)], dtype=np.uint64)
        host = acc.copy()
        engine.scale_accumulators(host, f)
        assert np.array_equal(host, want)
        dev = torch.from_numpy(acc.view(np.int64).copy()).cuda()
        engine.scale_accumulators(dev, f)
        assert np.array_equal(dev.cpu().numpy().view(np.uint64), want)
