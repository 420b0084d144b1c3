"""add.rn.bf16x2 (one rounding) against unpack -> fp32 add -> pack (two roundings), the form the chain epilogue used for
the up-sample add: bit-identical on every finite pair tried (CPU check, numpy)."""
import numpy as np


def to_bf16_bits(f32):
    u = f32.view(np.uint32)
    return ((u + (((u >> 16) & 1) + 0x7FFF)) >> 16).astype(np.uint16)


def bits_to_f32(b):
    return (b.astype(np.uint32) << 16).view(np.float32)


def bf16_from_f64(f):
    m, e = np.frexp(f)
    q = np.maximum(np.ldexp(1.0, e - 8), np.ldexp(1.0, -133))
    return (np.round(f / q) * q).astype(np.float32)       # np.round: half to even


def main(n=4_000_000, seed=0):
    rng = np.random.default_rng(seed)
    with np.errstate(over="ignore", invalid="ignore"):
        for label, a, b in (
            ("random bit patterns", bits_to_f32(rng.integers(0, 65536, n).astype(np.uint16)), bits_to_f32(rng.integers(0, 65536, n).astype(np.uint16))),
            ("activation-like", bits_to_f32(to_bf16_bits(rng.normal(0, 1, n).astype(np.float32))),
             bits_to_f32(to_bf16_bits((rng.normal(0, 1, n) * rng.choice([1e-3, 1, 30], n)).astype(np.float32)))),
        ):
            ok = np.isfinite(a) & np.isfinite(b)
            a, b = a[ok], b[ok]
            two = bits_to_f32(to_bf16_bits((a + b).astype(np.float32)))
            one = bf16_from_f64(a.astype(np.float64) + b.astype(np.float64))
            fin = np.isfinite(two) & np.isfinite(one) & (np.abs(one) < 3e38)
            bad = (two[fin] != one[fin]) & ~((two[fin] == 0) & (one[fin] == 0))
            print(f"{label}: {int(fin.sum())} pairs, {int(bad.sum())} differ")
            assert not bad.any()


if __name__ == "__main__":
    main()
