"""Running statistics: C oracle vs the reference text (CPU), CUDA kernel vs the oracle (GPU). Bar: bit-exact."""
import hashlib
import json
import os

import numpy as np
import pytest

from tests import helpers as H

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_stats.json")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_oracle_stats_match_the_golden_hashes(oracle_lib):
    got = H.stats_run(oracle_lib.OracleStats())
    assert [sha(a) for a in got] == json.load(open(GOLD))["arrays"]


def test_oracle_stats_equal_the_reference_text(oracle_lib):
    O = oracle_lib
    if not O.RefStats.available():
        pytest.skip("oracle/_ref/libluwref_stats.so is built only where /root/reference exists")
    for a, b in zip(H.stats_run(O.OracleStats()), H.stats_run(O.RefStats())):
        assert np.array_equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("Nx", [50, 64], ids=["padded-rows", "dense-rows"])
def test_cuda_stats_equal_oracle(oracle_lib, Nx):
    """luw_stats_accumulate on the device fields == the reference's host loop on the same samples, bit for bit; download in the layout of u / rho."""
    from latticeurbanwind_b200 import _cabi as A
    from latticeurbanwind_b200.domain import Domain, Stats
    O = oracle_lib
    assert H.STATS_N == 50 * 10 * 10
    Ny, Nz = (10, 10) if Nx == 50 else (10, 7)
    N = Nx * Ny * Nz
    with Domain(Nx, Ny, Nz, precision=1, features=0, w=1.0, arith=0) as d:
        st = Stats(d)
        orc = O.OracleStats()
        u_avg, rho_avg = np.zeros(3 * N, np.float32), np.zeros(N, np.float32)
        m2 = [np.zeros(N, np.float32) for _ in range(3)]
        for rho, u in H.stats_samples():
            rho, u = rho[:N].copy(), np.concatenate([u[c * H.STATS_N:c * H.STATS_N + N] for c in range(3)])
            d.rho[:], d.u[:] = rho, u
            d.write_to_device(A.FIELD_RHO); d.write_to_device(A.FIELD_U)
            st.accumulate()
            orc.accumulate(rho, u, u_avg, rho_avg, *m2)
        mean_u, m2_u, mean_rho, count = st.download()
        assert count == H.STATS_SAMPLES
        assert np.array_equal(mean_rho, rho_avg)
        for c in range(3):
            assert np.array_equal(mean_u[c * N:(c + 1) * N], u_avg[c::3])
            assert np.array_equal(m2_u[c * N:(c + 1) * N], m2[c])
        st.reset()
        st.accumulate()
        mean_u, m2_u, mean_rho, count = st.download()
        assert count == 1 and np.array_equal(mean_rho, d.rho) and np.all(m2_u == 0.0)
        st.close()
