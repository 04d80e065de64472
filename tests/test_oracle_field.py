"""Goldilocks arithmetic of the oracle against Python big integers (pinned by definition)."""
import random

P = 0xFFFFFFFF00000001


def test_field_ops_match_bigint(orc):
    rnd = random.Random(1)
    edge = [0, 1, 2, P - 1, P - 2, 0xFFFFFFFF, 0x100000000, 0xFFFFFFFF00000000, (1 << 63) % P]
    vals = edge + [rnd.randrange(P) for _ in range(300)]
    for a in vals[:40]:
        for b in vals:
            assert orc.orc_gl_mul(a, b) == a * b % P
            assert orc.orc_gl_add(a, b) == (a + b) % P
            assert orc.orc_gl_sub(a, b) == (a - b) % P
    for a in vals[1:60]:
        assert orc.orc_gl_mul(a, orc.orc_gl_inv(a)) == 1
