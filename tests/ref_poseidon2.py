"""Second, independent restatement of Poseidon2-Goldilocks-12 in pure Python big-int arithmetic,
used only to cross-check the C oracle (oracle/poseidon2.c).  Written from the published
construction (Grassi-Khovratovich-Schofnegger, 'Poseidon2'): explicit matrices, no add chains."""
import os
import sys

P = 0xFFFFFFFF00000001
M4 = [[5, 7, 1, 3], [4, 6, 1, 1], [1, 3, 5, 7], [1, 1, 4, 6]]
SHIFTS = [4, 14, 11, 8, 0, 5, 2, 9, 13, 6, 3, 12]


def constants():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import gen_poseidon2_constants as g
    return g.constants()


def external_matrix():
    m = [[0] * 12 for _ in range(12)]
    for bi in range(3):
        for bj in range(3):
            k = 2 if bi == bj else 1
            for i in range(4):
                for j in range(4):
                    m[4 * bi + i][4 * bj + j] = k * M4[i][j]
    return m


def inner_matrix():
    return [[(1 + (1 << SHIFTS[i])) if i == j else 1 for j in range(12)] for i in range(12)]


def matmul(m, s):
    return [sum(m[i][j] * s[j] for j in range(12)) % P for i in range(12)]


def permutation(state, rc=None):
    rc = rc or constants()
    me, mi = external_matrix(), inner_matrix()
    s = matmul(me, [x % P for x in state])
    for r in range(30):
        if r < 4 or r >= 26:
            s = [pow((s[i] + rc[12 * r + i]) % P, 7, P) for i in range(12)]
            s = matmul(me, s)
        else:
            s[0] = pow((s[0] + rc[12 * r]) % P, 7, P)
            s = matmul(mi, s)
    return s
