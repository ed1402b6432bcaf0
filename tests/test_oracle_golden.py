"""The oracle restatement (oracle/refpath.py) must reproduce the fixtures that the UNMODIFIED
reference produced in the authoring container (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import inputs, refpath, weights


def _sd_final():
    from realcamnet_b200 import raw2bit  # host mirror only provides names/shapes; no kernels run here

    m = raw2bit.raw_compression_tcm_final()
    weights.fill_(m, seed=0)
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


@pytest.fixture(scope="module")
def final_case(golden_dir):
    g = np.load(os.path.join(golden_dir, "final_T256.npz"))
    sd = _sd_final()
    x = inputs.make_inputs(256, seed=1234)
    return g, sd, x


def test_inputs_and_weights_reproduce(final_case):
    g, sd, x = final_case
    assert abs(float(x[0].double().sum()) - float(g["raw_sum"])) < 1e-6
    assert abs(float(x[1].double().sum()) - float(g["cond_sum"])) < 1e-6
    assert abs(weights.checksum(sd)["abs_sum"] - float(g["weights_abs_sum"])) < 1e-3


def test_final_forward_matches_reference_fixture(final_case):
    g, sd, x = final_case
    out = refpath.final_forward(sd, x)
    np.testing.assert_array_equal(out["y"].numpy(), g["y"])
    np.testing.assert_array_equal(out["para"]["means"].numpy(), g["means"])
    np.testing.assert_array_equal(out["para"]["scales"].numpy(), g["scales"])
    np.testing.assert_array_equal(out["likelihoods"]["y"].numpy(), g["lik_y"])
    np.testing.assert_array_equal(out["likelihoods"]["z"].numpy(), g["lik_z"])
    np.testing.assert_array_equal(out["lft"].numpy(), g["lft"])
    np.testing.assert_array_equal(out["lsc"][:, ::8, ::16, ::16].numpy(), g["lsc_sub"])
    np.testing.assert_array_equal(out["x_hat"][:, :, ::4, ::4].numpy(), g["x_hat_sub"])
    assert abs(float(out["x_hat"].double().abs().sum()) - float(g["x_hat_abs_sum"])) < 1e-6


def test_final_compress_decompress_match_reference_fixture(final_case):
    g, sd, x = final_case
    c = refpath.final_compress(sd, x)
    assert c["strings"][0][0] == g["y_string"].tobytes()
    assert c["strings"][1][0] == g["z_string"].tobytes()
    assert tuple(c["shape"]) == tuple(g["shape"])
    d = refpath.final_decompress(sd, c["strings"], c["shape"])
    np.testing.assert_array_equal(d["x_hat"][:, :, ::4, ::4].numpy(), g["dec_x_hat_sub"])


def test_entropy_tables_match_fixture(final_case):
    g, sd, _ = final_case
    gc = refpath._gc()
    np.testing.assert_array_equal(gc.quantized_cdf[[0, 1, 31, 63]].numpy(), g["gc_cdf_rows"])
    assert int(gc.quantized_cdf.long().sum()) == int(g["gc_cdf_sum"])
    np.testing.assert_array_equal(gc.cdf_length.numpy(), g["gc_cdf_length"])
    np.testing.assert_array_equal(gc.offset.numpy(), g["gc_offset"])
    eb = refpath._eb(sd)
    eb.update(force=True)
    np.testing.assert_array_equal(eb.quantized_cdf.numpy(), g["eb_cdf"])
    np.testing.assert_array_equal(eb.offset.numpy(), g["eb_offset"])


def test_liteisp_matches_reference_fixture(golden_dir):
    from realcamnet_b200 import LiteISP

    g = np.load(os.path.join(golden_dir, "liteisp_T256.npz"))
    m = LiteISP.LiteISPNet_GFM_LSC()
    weights.fill_(m, seed=0)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    assert abs(weights.checksum(sd)["abs_sum"] - float(g["weights_abs_sum"])) < 1e-3
    o = refpath.liteisp_gfm_lsc_forward(sd, inputs.make_inputs(256, seed=1235))
    np.testing.assert_array_equal(o[:, :, ::2, ::2].numpy(), g["out_sub"])


@pytest.mark.parametrize("dim", [80, 200])
def test_gma_block_matches_reference_fixture(golden_dir, dim):
    from realcamnet_b200 import groupmix

    g = np.load(os.path.join(golden_dir, f"gma_dim{dim}.npz"))
    m = groupmix.GMA_Block(dim, 8)
    weights.fill_(m, seed=0)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    o = refpath.gma_block(sd, torch.from_numpy(g["x"]), (24, 16), 8)
    np.testing.assert_allclose(o.numpy(), g["out"], rtol=0, atol=2e-6)


# ----------------------------------------------------------------------------- fixtures added later in round 1
def test_gma_wrappers_match_reference_fixture(golden_dir):
    """ConvGMABlock / GMAAtten in the reference's own test_gma configuration (raw2bit.py:4361-4367)."""
    from realcamnet_b200 import raw2bit

    g = np.load(os.path.join(golden_dir, "conv_gma_block.npz"))
    m = raw2bit.ConvGMABlock(64, 80, 10, drop_path=0.)
    weights.fill_(m, seed=0)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    o = refpath.conv_gma_block(sd, "", torch.from_numpy(g["x"]), 64, 8)
    np.testing.assert_allclose(o.numpy(), g["out"], rtol=0, atol=2e-6 * float(np.abs(g["out"]).max()))
    g = np.load(os.path.join(golden_dir, "gma_atten.npz"))
    m = raw2bit.GMAAtten(320, 320, 25, 0., 200)
    weights.fill_(m, seed=0)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    o = refpath.gma_atten(sd, "", torch.from_numpy(g["x"]), 8)
    np.testing.assert_allclose(o.numpy(), g["out"], rtol=0, atol=2e-6 * float(np.abs(g["out"]).max()))


def test_tcm_matches_reference_fixture(golden_dir):
    """TCM (tcm.py:320-637): forward, compress bytes and decompress of the oracle == the unmodified reference's."""
    from realcamnet_b200 import tcm

    g = np.load(os.path.join(golden_dir, "tcm_T256.npz"))
    m = tcm.TCM()
    weights.fill_(m, seed=0)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    assert abs(weights.checksum(sd)["abs_sum"] - float(g["weights_abs_sum"])) < 1e-3
    x = torch.rand(1, 3, 256, 256, generator=torch.Generator().manual_seed(1236))
    assert abs(float(x.double().sum()) - float(g["x_sum"])) < 1e-6
    out = refpath.tcm_forward(sd, x)
    np.testing.assert_array_equal(out["para"]["y"].numpy(), g["y"])
    np.testing.assert_array_equal(out["para"]["means"].numpy(), g["means"])
    np.testing.assert_array_equal(out["para"]["scales"].numpy(), g["scales"])
    np.testing.assert_array_equal(out["likelihoods"]["y"][:, ::4].numpy(), g["lik_y_sub"])
    np.testing.assert_array_equal(out["likelihoods"]["z"].numpy(), g["lik_z"])
    np.testing.assert_array_equal(out["x_hat"][:, :, ::2, ::2].numpy(), g["x_hat_sub"])
    c = refpath.tcm_compress(sd, x)
    assert c["strings"][0][0] == g["y_string"].tobytes() and c["strings"][1][0] == g["z_string"].tobytes()
    d = refpath.tcm_decompress(sd, c["strings"], c["shape"])
    np.testing.assert_array_equal(d["x_hat"][:, :, ::2, ::2].numpy(), g["dec_x_hat_sub"])


def test_liteisp_plain_matches_reference_fixture(golden_dir):
    """LiteISPNet (LiteISP.py:2322-2412), SURVEY 8f-4: the oracle restatement == the unmodified reference."""
    from realcamnet_b200 import LiteISP

    g = np.load(os.path.join(golden_dir, "liteisp_plain_T256.npz"))
    m = LiteISP.LiteISPNet()
    weights.fill_(m, seed=0)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    assert abs(weights.checksum(sd)["abs_sum"] - float(g["weights_abs_sum"])) < 1e-3
    o = refpath.liteisp_forward(sd, inputs.make_inputs(256, seed=1237))
    np.testing.assert_array_equal(o[:, :, ::2, ::2].numpy(), g["out_sub"])


@pytest.mark.parametrize("name,seed,fn", [("ISPUNet_GFM_LSC", 1241, "ispunet_gfm_lsc_forward"), ("ResUNet", 1242, "resunet_forward"),
                                          ("MWISP", 1243, "mwisp_forward")])
def test_isp_variants_match_reference_fixtures(golden_dir, name, seed, fn):
    """SURVEY 8f-4 (LiteISP.py:1228-1381, 2038-2146, 2149-2218): the oracle restatement == the unmodified reference, bit for bit."""
    from realcamnet_b200 import LiteISP

    g = np.load(os.path.join(golden_dir, f"isp_{name}_T128.npz"))
    m = getattr(LiteISP, name)()
    weights.fill_(m, seed=0)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    assert abs(weights.checksum(sd)["abs_sum"] - float(g["weights_abs_sum"])) < 1e-3
    o = getattr(refpath, fn)(sd, inputs.make_inputs(128, seed=seed, cond_size=128))
    np.testing.assert_array_equal(o[:, :, ::2, ::2].numpy(), g["out_sub"])
