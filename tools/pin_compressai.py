"""Pins the restated CompressAI subset (oracle/cai.py, oracle/rans.py) against upstream CompressAI -- the moment it is importable.

The reference imports CompressAI un-vendored and un-pinned (models/tcm.py:1-11, models/raw2bit.py:5-12); this image has no wheel
and no network, so every entropy-path parity claim of this repository reads "vs restated oracle" (DESIGN.md section 2).  This tool
closes that gap wherever `import compressai` succeeds:

  python tools/pin_compressai.py            -> prints one PASS/FAIL line per check, exit code 1 on any mismatch
                                               (exit code 3 = compressai not importable: nothing could be pinned)

Checks (all on the known-answer inputs of tests/test_rans_oracle.py / tests/test_abi_and_names.py):
  1. pmf_to_quantized_cdf: compressai._CXX.pmf_to_quantized_cdf == oracle.rans.pmf_to_quantized_cdf on random and degenerate pmfs
  2. GaussianConditional tables: update_scale_table(get_scale_table()) -> quantized_cdf / cdf_length / offset identical
  3. EntropyBottleneck(192) with the name-keyed weights: update() tables, eval forward (outputs, likelihoods), compress() bytes
  4. rANS: BufferedRansEncoder.encode_with_indexes + flush bytes == oracle.rans / rans_c.c bytes on a 50k-symbol sigma sweep
     with bypass symbols; RansDecoder decodes the oracle's stream
  5. GDN / ResidualBlockWithStride / ResidualBlockUpsample / AttentionBlock forward on seeded inputs (max abs diff)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch


def main():
    try:
        import compressai
        from compressai import ans as cans
        from compressai import entropy_models as cem
        from compressai import layers as clay
        from compressai._CXX import pmf_to_quantized_cdf as c_pmf
    except Exception as e:  # noqa: BLE001
        print(f"compressai is not importable here ({type(e).__name__}: {e}); parity stays 'unpinned vs upstream'")
        return 3
    from oracle import cai, rans, refpath
    from realcamnet_b200 import synthetic

    print("compressai", getattr(compressai, "__version__", "?"))
    bad = 0

    def check(name, ok, note=""):
        nonlocal bad
        print(("PASS " if ok else "FAIL ") + name + (" -- " + note if note else ""))
        bad += 0 if ok else 1

    # 1. pmf -> quantised CDF
    g = np.random.default_rng(0)
    ok = True
    for n in (2, 3, 17, 257, 3000):
        for _ in range(20):
            p = g.random(n).astype(np.float32) ** 8
            p /= p.sum()
            ok &= list(c_pmf(p.tolist(), 16)) == list(rans.pmf_to_quantized_cdf(p, 16))
    check("pmf_to_quantized_cdf", ok)

    # 2. Gaussian tables
    a, b = cem.GaussianConditional(None), cai.GaussianConditional(None)
    a.update_scale_table(cai.get_scale_table()), b.update_scale_table(cai.get_scale_table())
    check("GaussianConditional tables", torch.equal(a.quantized_cdf, b.quantized_cdf) and torch.equal(a.cdf_length, b.cdf_length) and
          torch.equal(a.offset, b.offset))

    # 3. EntropyBottleneck
    b = cai.EntropyBottleneck(192)
    synthetic.fill_(b, seed=5)
    a = cem.EntropyBottleneck(192)
    a.load_state_dict(b.state_dict())
    a.eval(), b.eval(), a.update(force=True), b.update(force=True)
    z = torch.randn(2, 192, 6, 5, generator=torch.Generator().manual_seed(3)) * 3
    (za, la), (zb, lb) = a(z), b(z)
    check("EntropyBottleneck tables", torch.equal(a.quantized_cdf, b.quantized_cdf) and torch.equal(a.offset, b.offset))
    check("EntropyBottleneck forward", torch.equal(za, zb), f"likelihood max abs diff {float((la - lb).abs().max()):.2e}")
    check("EntropyBottleneck compress bytes", a.compress(z) == b.compress(z))

    # 4. rANS
    gc = cai.GaussianConditional(None)
    gc.update_scale_table(cai.get_scale_table())
    cdf, sizes, offs = refpath._tables(gc)
    n = 50000
    sigma = np.exp(g.uniform(np.log(0.11), np.log(64), n)).astype(np.float32)
    sym = np.round(sigma * g.standard_normal(n)).astype(np.int32)
    esc = g.random(n) < 0.001
    sym[esc] = g.choice([-5000, 5000, -70000, 70000, 2 ** 30], size=int(esc.sum()))
    idx = gc.build_indexes(torch.from_numpy(sigma)).numpy().astype(np.int32)
    enc = cans.BufferedRansEncoder()
    enc.encode_with_indexes(sym.tolist(), idx.tolist(), cdf.tolist(), sizes.tolist(), offs.tolist())
    up = enc.flush()
    mine = refpath.encode_stream(sym, idx, gc)
    check("rANS encoder bytes", up == mine, f"{len(up)} vs {len(mine)} bytes")
    dec = cans.RansDecoder()
    dec.set_stream(mine)
    check("rANS decoder", dec.decode_stream(idx.tolist(), cdf.tolist(), sizes.tolist(), offs.tolist()) == sym.tolist())

    # 5. layers
    gen = torch.Generator().manual_seed(11)
    x = torch.randn(1, 32, 24, 20, generator=gen)
    for name, mk in (("GDN", lambda L: L.GDN(32)), ("IGDN", lambda L: L.GDN(32, inverse=True)),
                     ("ResidualBlockWithStride", lambda L: L.ResidualBlockWithStride(32, 32, 2)),
                     ("ResidualBlockUpsample", lambda L: L.ResidualBlockUpsample(32, 32, 2)),
                     ("ResidualBlock", lambda L: L.ResidualBlock(32, 32)), ("AttentionBlock", lambda L: L.AttentionBlock(32))):
        mb = mk(cai)
        synthetic.fill_(mb, seed=2)
        ma = mk(clay)
        ma.load_state_dict(mb.state_dict())
        d = float((ma.eval()(x) - mb.eval()(x)).abs().max())
        check(f"layers.{name}", d < 1e-6, f"max abs diff {d:.2e}")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
