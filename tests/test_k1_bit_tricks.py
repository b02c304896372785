"""CPU restatement of K1's integer-ALU number conversions (k1_stream.cuh: f32_f64, f32_raw, rnd_f32) in 64-bit integer
arithmetic, checked against the real float <-> double conversions.  This pins the MATH of the instruction sequences
(IMAD.WIDE field shift, exponent re-bias, LEA carry rounding); the GPU parity tests pin the code."""
import numpy as np

M32 = np.uint64(0xFFFFFFFF)


def f32_bits(f):
    return f.astype(np.float32).view(np.uint32).astype(np.uint64)


def as_double(hi, lo):
    return ((hi & M32) << np.uint64(32) | (lo & M32)).astype(np.uint64).view(np.float64)


def f32_f64(u):                       # mul.wide.u32 by 2^29, then hi += (sign | 0x38000000)
    w = (u & np.uint64(0x7FFFFFFF)) * np.uint64(0x20000000)
    hi = (w >> np.uint64(32)) + ((u & np.uint64(0x80000000)) | np.uint64(0x38000000))
    return as_double(hi, w)


def f32_raw(u):                       # same without the re-bias: value * 2^-896
    w = (u & np.uint64(0x7FFFFFFF)) * np.uint64(0x20000000)
    hi = (w >> np.uint64(32)) | (u & np.uint64(0x80000000))
    return as_double(hi, w)


def rnd_f32(x):                       # shl 2 / add.cc 0x80000000 -> carry = bit 29; addc.cc lo + 0x0FFFFFFF; addc hi; mask
    b = x.view(np.uint64)
    lo, hi = b & M32, b >> np.uint64(32)
    carry = ((lo << np.uint64(2)) & M32) >> np.uint64(31)            # bit 29 of lo
    s = lo + np.uint64(0x0FFFFFFF) + carry
    lo2, hi2 = s & M32, hi + (s >> np.uint64(32))
    return as_double(hi2, lo2 & np.uint64(0xE0000000))


def samples(n=400_000, seed=1):
    rng = np.random.default_rng(seed)
    mant = rng.uniform(1.0, 2.0, n)
    expo = rng.integers(-60, 60, n)
    sign = rng.choice([-1.0, 1.0], n)
    return sign * mant * np.exp2(expo)


def test_float_to_double_on_the_integer_alu_is_exact():
    f = samples().astype(np.float32)
    f = np.concatenate([f, np.array([1.0, -1.0, 0.5, 3.4028235e38, -3.4028235e38, 1.1754944e-38], np.float32)])
    u = f32_bits(f)
    assert np.array_equal(f32_f64(u), f.astype(np.float64))
    assert np.array_equal(f32_raw(u) * np.exp2(896.0), f.astype(np.float64))     # power-of-two scaling: exact
    # +-0 stays +-0 in the raw form (the re-biased form maps it to 2^-127-sized values, documented in k1_reduce.cuh)
    z = f32_raw(f32_bits(np.array([0.0, -0.0], np.float32)))
    assert z[0] == 0.0 and z[1] == 0.0 and np.signbit(z[1])


def test_float32_round_trip_on_the_integer_alu_is_round_to_nearest_even():
    x = samples(seed=2)
    assert np.array_equal(rnd_f32(x), x.astype(np.float32).astype(np.float64))
    # exact ties (bit 28 set, lower bits clear): to even, both ways
    base = np.float64(1.0) + np.exp2(-23.0) * np.arange(0, 64)                     # float32 grid points near 1
    ties = base + np.exp2(-24.0)
    assert np.array_equal(rnd_f32(ties), ties.astype(np.float32).astype(np.float64))
    # mantissa overflow into the exponent
    y = np.array([np.nextafter(2.0, 0.0), -np.nextafter(4.0, 0.0), 0.0], np.float64)
    assert np.array_equal(rnd_f32(y), y.astype(np.float32).astype(np.float64))
