"""Host-side arithmetic of the fp16 operand form (no GPU): the Python mirror of csrc/common.cuh::f16_scale_of and the
precision the hi / lo pair keeps, checked on a bit-level model of what agnn_split_f16 computes."""
import struct

import pytest
import torch

from analysisgnn_b200 import linalg


def scale_of(amax: float) -> float:
    """csrc/common.cuh::f16_scale_of, restated on the bit pattern."""
    bits = struct.unpack("<I", struct.pack("<f", amax))[0] & 0x7FFFFFFF
    e = bits >> 23
    if e == 0 or e == 255:
        return 1.0
    k = max(-100, min(100, 13 - (e - 127)))
    return 2.0 ** k


@pytest.mark.parametrize("amax", [1.0, 2.0, 3.999, 4.0, 0.5, 1e-30, 1e30, 8192.0, 16383.9, 0.0, float("inf"), 1e-39,
                                  1.17e-38, 65504.0, 3e-6, 7e4])
def test_python_scale_matches_the_kernel_rule(amax):
    got = float(linalg.f16_scale(torch.tensor([amax], dtype=torch.float32)))
    assert got == scale_of(amax)
    if 1e-25 < amax < 1e25:
        assert 2.0 ** 13 <= amax * got < 2.0 ** 14


@pytest.mark.parametrize("magnitude", [1.0, 3e-6, 7e4])
def test_pair_keeps_22_bits_down_to_2_pow_minus_17_of_the_amax(magnitude):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(300, 200, generator=g) * magnitude
    x[::3] *= 2.0 ** -16                                # rows far below the amax
    s = scale_of(float(x.abs().max()))
    xs = x * s
    hi = xs.half()
    lo = (xs - hi.float()).half()
    back = (hi.double() + lo.double()) / s
    err = (back - x.double()).abs()
    rel = err / x.double().abs().clamp_min(1e-300)
    big = x.abs() >= float(x.abs().max()) * 2.0 ** -17
    assert float(rel[big].max()) <= 2.0 ** -21          # two 11-bit significands
    assert float(err.max()) <= max(float((x.double().abs() * 2.0 ** -21).max()), 2.0 ** -24 / s)
    assert not torch.isinf(hi).any() and float(hi.abs().max()) < 2.0 ** 14


def test_parity_operand_switch():
    old = linalg.parity_operands()
    try:
        linalg.set_parity_operands("tf32")
        assert linalg.parity_operands() == "tf32"
        linalg.set_parity_operands("f16")
        assert linalg.parity_operands() == "f16"
        with pytest.raises(ValueError):
            linalg.set_parity_operands("fp8")
    finally:
        linalg.set_parity_operands(old)


def test_no_cpu_path_for_the_operand_producers():
    from analysisgnn_b200 import _lib
    with pytest.raises(_lib.AgnnError):
        linalg.split_f16(torch.randn(8, 8))
    with pytest.raises(_lib.AgnnError):
        linalg.amax_into(torch.zeros(1), torch.randn(8, 8))


def test_split_cache_is_bounded_without_begin_step():
    """Callers that never call ``begin_step`` (a plain torch optimizer loop) must not grow the weight-split cache
    without bound: the oldest half goes when the cap is passed."""
    linalg.begin_step()
    for i in range(linalg._SPLIT_CACHE_MAX_ENTRIES + 10):
        linalg._remember(("k", i), i)
    assert len(linalg._split_cache) <= linalg._SPLIT_CACHE_MAX_ENTRIES
    assert ("k", linalg._SPLIT_CACHE_MAX_ENTRIES + 9) in linalg._split_cache and ("k", 0) not in linalg._split_cache
    linalg.begin_step()
    assert not linalg._split_cache


def test_there_is_no_backend_switch_and_strict_mode_raises_on_library_routes():
    """north_star: no multi-backend dispatch.  The GEMM has one implementation; routes off the hand-written kernels
    are counted and, in strict mode (bench.py, full-size parity tests), refused."""
    from analysisgnn_b200 import _lib
    assert not hasattr(linalg, "set_backend") and not hasattr(linalg, "backend")
    import inspect
    src = inspect.getsource(linalg)
    assert "torch.mm(" not in src and "addmm" not in src
    old = _lib.strict()
    try:
        _lib.set_strict(False)
        before = _lib.library_routes.get("probe", 0)
        _lib.library_route("probe")
        assert _lib.library_routes["probe"] == before + 1
        _lib.set_strict(True)
        with pytest.raises(_lib.AgnnError):
            _lib.library_route("probe")
    finally:
        _lib.set_strict(old)


def test_gemm_work_item_counter_protocol():
    """The hand-out protocol of gemm_kernel's work-item counter (csrc/gemm.cu, GemmGroup::sched), replayed on the host
    under random interleavings: every CTA draws ids with fetch-and-add until it gets one >= n_items, the CTA that draws
    n_items + grid - 1 (the last draw of the launch) writes zero.  Whatever the order -- CTAs that start late, CTAs
    that never get an item -- every item is handed out exactly once and the counter is zero afterwards, which is what
    lets the next launch reuse it without a memset."""
    import random
    for trial in range(200):
        rng = random.Random(trial)
        n_items, grid = rng.randint(1, 400), rng.randint(1, 148)
        counter = [0]
        got, done = [[] for _ in range(grid)], [False] * grid
        started = [False] * grid
        resets = 0
        while not all(done):
            cta = rng.choice([c for c in range(grid) if not done[c]])
            if not started[cta] and rng.random() < 0.7 and any(started):      # some CTAs wait for an SM for a long time
                continue
            started[cta] = True
            drawn = counter[0]
            counter[0] += 1                                   # atomicAdd
            if drawn < n_items:
                got[cta].append(drawn)
                continue
            if drawn == n_items + grid - 1:
                counter[0] = 0                                # atomicExch by the last draw
                resets += 1
            done[cta] = True
        assert sorted(i for g in got for i in g) == list(range(n_items))
        assert counter[0] == 0 and resets == 1
