"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol declared in
include/rcn_b200.h, and the host mirrors expose the reference's parameter names/shapes."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "rcn_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rcn_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from realcamnet_b200 import _C

    lib = _C.lib()
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in rcn_b200.h but not exported"
        assert n in _C.PROTOTYPES, f"{n} has no ctypes prototype"
    assert lib.rcn_version() >= 100


def test_host_range_coder_matches_oracle_bytes():
    """The native C++ coder is host code: its bytes can be checked against the oracle without a GPU."""
    import numpy as np

    from oracle import cai, rans, refpath
    from realcamnet_b200 import entropy_models as em

    gc = cai.GaussianConditional(None)
    gc.update_scale_table(cai.get_scale_table())
    cdf, sizes, offs = refpath._tables(gc)
    g = np.random.default_rng(5)
    n = 50000
    sigma = np.exp(g.uniform(np.log(0.11), np.log(64), n)).astype(np.float32)
    sym = np.round(sigma * g.standard_normal(n)).astype(np.int32)
    esc = g.random(n) < 0.001
    sym[esc] = g.choice([-5000, 5000, -70000, 70000, 2 ** 30], size=int(esc.sum()))
    idx = gc.build_indexes(torch.from_numpy(sigma)).numpy().astype(np.int32)
    ours = em.rans_encode(sym, idx, cdf, sizes, offs)
    assert ours == refpath.encode_stream(sym, idx, gc)
    assert ours[:4096] == rans.encode_with_indexes(sym, idx, cdf.tolist(), sizes, offs)[:4096]
    d = em.RansDecoder()
    d.set_stream(ours)
    a = d.decode_stream(idx[:1234], cdf, sizes, offs)
    b = d.decode_stream(idx[1234:], cdf, sizes, offs)
    assert np.array_equal(np.concatenate([a, b]), sym)
    assert em.rans_encode(np.zeros(0, np.int32), np.zeros(0, np.int32), cdf, sizes, offs) == bytes([0, 0, 0, 0x80, 0, 0, 0, 0])


def test_host_decoder_rejects_corrupt_input():
    """The decoder treats the stream and the indexes as untrusted: out-of-range table rows, truncated streams and
    impossible bypass lengths return an error instead of reading out of bounds."""
    import numpy as np

    from oracle import cai, refpath
    from realcamnet_b200 import entropy_models as em

    gc = cai.GaussianConditional(None)
    gc.update_scale_table(cai.get_scale_table())
    cdf, sizes, offs = refpath._tables(gc)
    g = np.random.default_rng(11)
    n = 4000
    idx = g.integers(0, 64, n).astype(np.int32)
    sym = np.round(g.standard_normal(n) * 3).astype(np.int32)
    sym[::97] = 40000                                              # escapes
    stream = em.rans_encode(sym, idx, cdf, sizes, offs)
    for bad in (np.array([64], np.int32), np.array([-1], np.int32), np.array([1 << 20], np.int32)):
        d = em.RansDecoder()
        d.set_stream(stream)
        with pytest.raises(RuntimeError, match="outside"):
            d.decode_stream(bad, cdf, sizes, offs)
    d = em.RansDecoder()
    d.set_stream(stream[:len(stream) // 2 // 4 * 4])                # truncated: the decoder runs off the end
    with pytest.raises(RuntimeError, match="truncated|corrupt"):
        d.decode_stream(idx, cdf, sizes, offs)
    # random garbage decodes to *something* or errors out, but never crashes / hangs
    junk = g.integers(0, 256, 4096, dtype=np.uint8).tobytes()
    d = em.RansDecoder()
    d.set_stream(junk)
    try:
        d.decode_stream(idx, cdf, sizes, offs)
    except RuntimeError:
        pass
    d = em.RansDecoder()
    d.set_stream(stream)
    assert np.array_equal(d.decode_stream(idx, cdf, sizes, offs), sym)


def test_native_cdf_quantiser_matches_oracle():
    import numpy as np

    from oracle import rans
    from realcamnet_b200 import entropy_models as em

    g = np.random.default_rng(0)
    for n in (3, 17, 200, 3131):
        p = g.random(n).astype(np.float32) ** 8
        p[g.random(n) < 0.3] = 0
        p[0] = max(p[0], 1e-3)
        p /= p.sum()
        assert em.pmf_to_quantized_cdf(p).tolist() == rans.pmf_to_quantized_cdf(p)


def test_update_builds_the_oracle_tables():
    from oracle import cai, weights
    from realcamnet_b200 import entropy_models as em
    from realcamnet_b200.tcm import get_scale_table

    gc = em.GaussianConditional(None)
    gc.update_scale_table(get_scale_table())
    ref = cai.GaussianConditional(None)
    ref.update_scale_table(cai.get_scale_table())
    assert torch.equal(gc.quantized_cdf, ref.quantized_cdf)
    assert torch.equal(gc.cdf_length, ref.cdf_length) and torch.equal(gc.offset, ref.offset)
    eb, reb = em.EntropyBottleneck(192), cai.EntropyBottleneck(192)
    weights.fill_(eb, seed=3)
    reb.load_state_dict(eb.state_dict())
    eb.update(), reb.update()
    assert torch.equal(eb.quantized_cdf, reb.quantized_cdf) and torch.equal(eb.offset, reb.offset)


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="reference only exists in the authoring container")
@pytest.mark.parametrize("which", ["final", "liteisp", "gma80", "gma200", "tcm", "convgma", "gmaatten", "gmablock", "liteisp_plain",
                                   "ISPUNet_GFM_LSC", "ResUNet", "MWISP"])
def test_state_dict_names_match_reference(which):
    from oracle import ref_import

    ref = ref_import.import_reference()
    from realcamnet_b200 import LiteISP, groupmix, raw2bit, tcm

    if which in ("ISPUNet_GFM_LSC", "ResUNet", "MWISP"):
        a, b = getattr(ref.LiteISP, which)(), getattr(LiteISP, which)()
    elif which == "liteisp_plain":
        a, b = ref.LiteISP.LiteISPNet(), LiteISP.LiteISPNet()
    elif which == "tcm":
        a, b = ref.tcm.TCM(), tcm.TCM()
    elif which == "convgma":
        a, b = ref.raw2bit.ConvGMABlock(64, 80, 10, drop_path=0.), raw2bit.ConvGMABlock(64, 80, 10, drop_path=0.)
    elif which == "gmaatten":
        a, b = ref.raw2bit.GMAAtten(320, 320, 25, 0., 200), raw2bit.GMAAtten(320, 320, 25, 0., 200)
    elif which == "gmablock":
        a, b = ref.raw2bit.GMABlock(200, 25, 0.), raw2bit.GMABlock(200, 25, 0.)
    elif which == "final":
        a, b = ref.raw2bit.raw_compression_tcm_final(), raw2bit.raw_compression_tcm_final()
    elif which == "liteisp":
        a, b = ref.LiteISP.LiteISPNet_GFM_LSC(), LiteISP.LiteISPNet_GFM_LSC()
    else:
        dim = int(which[3:])
        a, b = ref.groupmix.GMA_Block(dim, 8), groupmix.GMA_Block(dim, 8)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa.keys()) == list(sb.keys())
    for k in sa:
        assert tuple(sa[k].shape) == tuple(sb[k].shape), k
    b.load_state_dict(sa)  # a reference checkpoint loads into the mirror


def test_product_fails_loudly_without_library(monkeypatch):
    from realcamnet_b200 import _C

    monkeypatch.setattr(_C, "_lib", None)
    monkeypatch.setattr(_C, "LIB_PATH", "/nonexistent/librcn_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _C.lib()


def test_packed_state_chain_matches_full_encoder():
    """rcn_rans_encode_packed (GPU front end + host state chain) must emit the bytes of the full host coder."""
    import numpy as np

    from oracle import cai, refpath
    from realcamnet_b200 import entropy_models as em

    gc = cai.GaussianConditional(None)
    gc.update_scale_table(cai.get_scale_table())
    cdf, sizes, offs = refpath._tables(gc)
    g = np.random.default_rng(11)
    n = 40000
    sigma = np.exp(g.uniform(np.log(0.11), np.log(200), n)).astype(np.float32)
    sym = np.round(sigma * g.standard_normal(n)).astype(np.int32)
    esc = g.random(n) < 0.002
    sym[esc] = g.choice([-5000, 5000, -70000, 70000, 2 ** 30], size=int(esc.sum()))
    idx = gc.build_indexes(torch.from_numpy(sigma)).numpy().astype(np.int32)
    # numpy emulation of the kernel's front end
    sentinel = sizes[idx] - 2
    v = sym.astype(np.int64) - offs[idx]
    raw = np.where(v < 0, -2 * v - 1, np.where(v >= sentinel, 2 * (v - sentinel), 0)).astype(np.uint32)
    escaped = (v < 0) | (v >= sentinel)
    b = np.where(escaped, sentinel, v).astype(np.int64)
    start = cdf[idx, b].astype(np.uint32)
    freq = (cdf[idx, b + 1] - cdf[idx, b]).astype(np.uint32)
    packed = (start << 16) | ((freq - 1) & 0xFFFF)
    raw[~escaped] = 0xDEADBEEF              # the kernel leaves raw[] undefined where the flag is clear
    ours = em.rans_encode_packed(packed, raw, escaped.astype(np.uint8))
    assert ours == em.rans_encode(sym, idx, cdf, sizes, offs) == refpath.encode_stream(sym, idx, gc)
