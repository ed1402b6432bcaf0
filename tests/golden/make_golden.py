"""Generates tests/golden/*.npz from the UNMODIFIED reference (authoring container only).

Run:  python tests/golden/make_golden.py
It imports /root/reference/models/*.py through oracle/ref_import.py (stub modules + the restated
CompressAI subset), fills the reference modules with the name-keyed deterministic weights of
oracle/weights.py, runs them on the seeded inputs of oracle/inputs.py and stores the outputs.
The GPU box has no /root/reference: there the fixtures pin the oracle restatement
(tests/test_oracle_golden.py) which in turn checks the CUDA product.

Provenance of every array: produced by reference code (models/raw2bit.py, LiteISP.py, groupmix.py)
-- except the entropy-coder bytes/tables, which come from the restated CompressAI subset the
reference was run against ("parity unpinned" vs upstream CompressAI, see oracle/cai.py).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import inputs, ref_import, weights  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main(only_new=False, only_r2=False):
    ref = ref_import.import_reference()
    torch.set_grad_enabled(False)
    if only_r2:       # fixtures added in round 2: leave the committed earlier ones untouched
        return isp_variants(ref)
    if only_new:      # fixtures added later in round 1
        return extra(ref)
    base(ref)
    extra(ref)
    isp_variants(ref)


ISP_VARIANTS = (("ISPUNet_GFM_LSC", 1241, 128), ("ResUNet", 1242, 128), ("MWISP", 1243, 128))


def isp_variants(ref):
    """the remaining ISP variants (LiteISP.py:1228-1381, 2038-2146, 2149-2218), one 4x128x128 tile each."""
    for name, seed, T in ISP_VARIANTS:
        m = getattr(ref.LiteISP, name)().eval()
        weights.fill_(m, seed=0)
        x = inputs.make_inputs(T, seed=seed, cond_size=128)
        o = m(x)
        np.savez_compressed(os.path.join(OUT, f"isp_{name}_T{T}.npz"), out_sub=o[:, :, ::2, ::2].numpy(),
                            out_abs_sum=np.float64(o.double().abs().sum()),
                            weights_abs_sum=np.float64(weights.checksum(m.state_dict())["abs_sum"]))
        print(name, tuple(o.shape), float(o.abs().max()))


def base(ref):

    # ---- raw_compression_tcm_final, T=256 (raw2bit.py:1614-2027)
    m = ref.raw2bit.raw_compression_tcm_final().eval()
    weights.fill_(m, seed=0)
    x = inputs.make_inputs(256, seed=1234)
    out = m(x)
    m.update(force=True)
    c = m.compress(x)
    d = m.decompress(c["strings"], c["shape"])
    gc, eb = m.gaussian_conditional, m.entropy_bottleneck
    np.savez_compressed(
        os.path.join(OUT, "final_T256.npz"),
        weights_abs_sum=np.float64(weights.checksum(m.state_dict())["abs_sum"]),
        raw_sum=np.float64(x[0].double().sum()), cond_sum=np.float64(x[1].double().sum()),
        y=out["y"].numpy(), means=out["para"]["means"].numpy(), scales=out["para"]["scales"].numpy(),
        lik_y=out["likelihoods"]["y"].numpy(), lik_z=out["likelihoods"]["z"].numpy(),
        lft=out["lft"].numpy(), lsc_sub=out["lsc"][:, ::8, ::16, ::16].numpy(),
        x_hat_sub=out["x_hat"][:, :, ::4, ::4].numpy(), x_hat_abs_sum=np.float64(out["x_hat"].double().abs().sum()),
        dec_x_hat_sub=d["x_hat"][:, :, ::4, ::4].numpy(),
        y_string=np.frombuffer(c["strings"][0][0], dtype=np.uint8),
        z_string=np.frombuffer(c["strings"][1][0], dtype=np.uint8), shape=np.asarray(c["shape"]),
        gc_cdf_rows=gc.quantized_cdf[[0, 1, 31, 63]].numpy(), gc_cdf_sum=np.int64(gc.quantized_cdf.long().sum()),
        gc_cdf_length=gc.cdf_length.numpy(), gc_offset=gc.offset.numpy(), scale_table=gc.scale_table.numpy(),
        eb_cdf=eb.quantized_cdf.numpy(), eb_cdf_length=eb.cdf_length.numpy(), eb_offset=eb.offset.numpy(),
    )
    print("final_T256: y bytes", len(c["strings"][0][0]), "z bytes", len(c["strings"][1][0]))

    # ---- LiteISPNet_GFM_LSC, BASELINE config 1 (LiteISP.py:1924-2035)
    m = ref.LiteISP.LiteISPNet_GFM_LSC().eval()
    weights.fill_(m, seed=0)
    x = inputs.make_inputs(256, seed=1235)
    o = m(x)
    np.savez_compressed(os.path.join(OUT, "liteisp_T256.npz"), out_sub=o[:, :, ::2, ::2].numpy(),
                        out_abs_sum=np.float64(o.double().abs().sum()),
                        weights_abs_sum=np.float64(weights.checksum(m.state_dict())["abs_sum"]))

    # ---- GMA_Block, the two dims the reference instantiates (raw2bit.py:4362-4363)
    for dim in (80, 200):
        g = ref.groupmix.GMA_Block(dim, 8).eval()
        weights.fill_(g, seed=0)
        gen = torch.Generator().manual_seed(77 + dim)
        xx = torch.randn(2, 24 * 16, dim, generator=gen)
        o = g(xx, (24, 16))
        np.savez_compressed(os.path.join(OUT, f"gma_dim{dim}.npz"), x=xx.numpy(), out=o.numpy())


def extra(ref):
    # ---- GroupMix drop-in wrappers, the two instantiations of the reference's test_gma (raw2bit.py:4361-4367)
    g = ref.raw2bit.ConvGMABlock(64, 80, 10, drop_path=0.).eval()
    weights.fill_(g, seed=0)
    xx = torch.randn(1, 144, 32, 32, generator=torch.Generator().manual_seed(901))
    np.savez_compressed(os.path.join(OUT, "conv_gma_block.npz"), x=xx.numpy(), out=g(xx).numpy())
    g = ref.raw2bit.GMAAtten(320, 320, 25, 0., 200).eval()
    weights.fill_(g, seed=0)
    xx = torch.randn(1, 320, 32, 32, generator=torch.Generator().manual_seed(902))
    np.savez_compressed(os.path.join(OUT, "gma_atten.npz"), x=xx.numpy(), out=g(xx).numpy())

    # ---- LiteISPNet, the plain UNet variant (LiteISP.py:2322-2412)
    m = ref.LiteISP.LiteISPNet().eval()
    weights.fill_(m, seed=0)
    x = inputs.make_inputs(256, seed=1237)
    o = m(x)
    np.savez_compressed(os.path.join(OUT, "liteisp_plain_T256.npz"), out_sub=o[:, :, ::2, ::2].numpy(),
                        out_abs_sum=np.float64(o.double().abs().sum()),
                        weights_abs_sum=np.float64(weights.checksum(m.state_dict())["abs_sum"]))

    # ---- TCM, the RGB baseline with the same entropy model (tcm.py:320-637), 256x256
    m = ref.tcm.TCM().eval()
    weights.fill_(m, seed=0)
    x = torch.rand(1, 3, 256, 256, generator=torch.Generator().manual_seed(1236))
    out = m(x)
    m.update(force=True)
    c = m.compress(x)
    d = m.decompress(c["strings"], c["shape"])
    np.savez_compressed(
        os.path.join(OUT, "tcm_T256.npz"), x_sum=np.float64(x.double().sum()),
        weights_abs_sum=np.float64(weights.checksum(m.state_dict())["abs_sum"]),
        y=out["para"]["y"].numpy(), means=out["para"]["means"].numpy(), scales=out["para"]["scales"].numpy(),
        lik_y_sub=out["likelihoods"]["y"][:, ::4].numpy(), lik_z=out["likelihoods"]["z"].numpy(),
        x_hat_sub=out["x_hat"][:, :, ::2, ::2].numpy(), x_hat_abs_sum=np.float64(out["x_hat"].double().abs().sum()),
        dec_x_hat_sub=d["x_hat"][:, :, ::2, ::2].numpy(),
        y_string=np.frombuffer(c["strings"][0][0], dtype=np.uint8),
        z_string=np.frombuffer(c["strings"][1][0], dtype=np.uint8), shape=np.asarray(c["shape"]))
    print("tcm_T256: y bytes", len(c["strings"][0][0]), "z bytes", len(c["strings"][1][0]))
    print("done")


if __name__ == "__main__":
    main(only_new="--extra" in sys.argv, only_r2="--r2" in sys.argv)
