"""GPU: the numerics contract behind the slab test. RayIntersectsBox (reference src/Raytracer.cc:135-136) needs
correctly rounded quotients; the kernels compute them with a per-ray refined reciprocal + 3 FMAs, which must be
bit-identical to the IEEE divide over the whole operand domain where that path is taken."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_shared_reciprocal_divide_equals_ieee_divide(gpu, seed):
    bad, first = gpu.selftest_division(samples=1 << 33, seed=seed)
    assert bad == 0, f"{bad} mismatches; first: a={first[0]!r} d={first[1]!r} a/d={first[2]!r} fast={first[3]!r}"
