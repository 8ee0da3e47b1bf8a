"""Wire format (include/anchored_fusion.h): 4 L bits per pair cross PCIe, the tiles are rebuilt on the GPU.  The device
expansion must give back the packed tiles bit for bit, and the pipeline must deliver the same records either way."""
import ctypes

import numpy as np
import pytest

from conftest import hits_equal

pytestmark = pytest.mark.gpu


def _ragged_batch(af, rng, L, n, pad_byte):
    lens1 = rng.integers(max(1, L - 30), L + 1, n)
    lens2 = rng.integers(max(1, L - 30), L + 1, n)
    lens1[0] = L
    draw = lambda l: "".join("ACGTN"[c] for c in rng.choice(5, l, p=[.2475, .2475, .2475, .2475, .01]))
    return af.pack_pairs([draw(l) for l in lens1], [draw(l) for l in lens2], max_read_len=L, pad_byte=pad_byte)


@pytest.mark.parametrize("L,n", [(150, 1000), (36, 33), (101, 4097), (250, 64), (256, 31), (16, 40), (17, 1), (257, 5), (300, 100), (301, 64), (512, 33)])
def test_device_expansion_restores_the_tiles_bit_for_bit(L, n):
    import torch
    import anchored_fusion_b200 as af
    from anchored_fusion_b200._lib import check, lib
    rng = np.random.default_rng(L * 1000 + n)
    b = _ragged_batch(af, rng, L, n, 0x6C)
    packed = np.asarray(b.packed).view(np.uint32)
    wire = af.wire_from_packed(packed, L, n)
    assert wire.nbytes == ((n + 31) // 32) * ((4 * L + 31) // 32) * 128
    assert np.array_equal(af.wire_to_packed(wire, L, n, 0x6C), packed)                 # host twin
    d_wire = torch.from_numpy(wire.view(np.int32)).cuda()
    d_out = torch.full((packed.size + 64,), -1, dtype=torch.int32, device="cuda")      # 256 guard bytes behind the tiles
    check(lib().af_wire_expand_device(d_wire.data_ptr(), L, n, 0x6C, d_out.data_ptr(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    got = d_out.cpu().numpy().view(np.uint32)
    assert np.array_equal(got[: packed.size], packed)
    assert (got[packed.size:] == 0xFFFFFFFF).all()


def test_wire_pipeline_delivers_the_records_of_the_tile_pipeline_and_the_oracle():
    import anchored_fusion_b200 as af
    from anchored_fusion_b200._lib import lib
    from oracle import oracle
    spec = af.synth_spec(seed=4, ref_len=400_000, anchor_start=50_000, anchor_len=6783, read_len=150, sub_ppm=10_000, fusion_ppm=20_000)
    index = af.AnchorIndex(af.synth_anchor(spec))
    eng = af.Anchorer(index, 0)
    n = 200_000
    dev = af.synth_pairs_device(spec, 0, n, index.pad_byte, 0)
    tiles = dev.packed.cpu().numpy().view(np.uint32)
    want, _ = eng.anchor(dev)
    reads = oracle.synth_reads(oracle.as_synth(spec), 0, n, threads=4)
    ora = oracle.anchor_reads(oracle.encode(af.synth_anchor(spec)), reads, threads=4)
    assert len(ora) > 1000 and hits_equal(want, ora)
    wire = af.wire_from_packed(tiles, 150, n)
    assert wire.nbytes * 20 == tiles.nbytes * 19                                      # 76 bytes per pair instead of 80
    before = lib().af_pipeline_h2d_bytes(eng.pipeline(150, 32_768, 3))
    got, _ = eng.anchor_host(af.PackedBatch(wire, n, 150, 150), slot_pairs=32_768, n_slots=3, wire=True)
    assert hits_equal(got, want)
    assert lib().af_pipeline_h2d_bytes(eng.pipeline(150, 32_768, 3)) - before == wire.nbytes
    got2, _ = eng.anchor_host(af.PackedBatch(tiles, n, 150, 150), slot_pairs=32_768, n_slots=3)
    assert hits_equal(got2, want)


def test_wire_pipeline_with_ragged_reads_and_n():
    import anchored_fusion_b200 as af
    rng = np.random.default_rng(12)
    anchor = "".join("ACGT"[c] for c in rng.integers(0, 4, 3000))
    index = af.AnchorIndex(anchor)
    eng = af.Anchorer(index, 0)
    n, L = 3000, 101
    s1, s2 = [], []
    for i in range(n):
        for dst in (s1, s2):
            l = int(rng.integers(19, L + 1))
            if i % 3:
                at = int(rng.integers(0, len(anchor) - l))
                r = list(anchor[at:at + l])
                for _ in range(int(rng.integers(0, 3))):
                    r[int(rng.integers(0, l))] = "ACGTN"[int(rng.integers(0, 5))]
                r = "".join(r)
            else:
                r = "".join("ACGT"[c] for c in rng.integers(0, 4, l))
            dst.append(r)
    host = af.pack_pairs(s1, s2, max_read_len=L, pad_byte=index.pad_byte)
    want, _ = eng.anchor_host(host, slot_pairs=512, n_slots=2)
    assert len(want) > 1000 and host.uniform_len == 0 and host.n_nreads > 50
    wire = af.wire_from_packed(np.asarray(host.packed).view(np.uint32), L, n)
    hw = af.PackedBatch(wire, n, L, 0, lens=host.lens, nread_ids=host.nread_ids, nmask=host.nmask)
    got, _ = eng.anchor_host(hw, slot_pairs=512, n_slots=2, wire=True)
    assert hits_equal(got, want)
