"""Compressed-layer format (neural_network_compression_b200/common/storage.py): bit-exact
round trips of both layouts, with and without Huffman coding.  The CPU tests build the layer from the oracle's fit (the
format code is host-side integer work); the GPU test goes through utility.compress_weight and the device bit packer."""
import io

import numpy as np
import pytest

from neural_network_compression_b200.common import storage as S
from oracle import oracle as O
from . import _data as D


class _Fit:  # the fields of utility.KMeansResult the format reads
    def __init__(self, centers, labels, bits):
        self.cluster_centers_ = centers
        self.code_bits = bits
        self.packed_codes = O.pack_codes(labels, bits)


def _lenet_layer(bits=4):
    w = D.lenet5_tensors()[2][1].copy()  # dense 2450 x 256
    mask = O.prune_weigth(w, 1)
    space = O.init_centroids(w, bits, "linear")
    km = O.kmeans1d(w, space, mode=O.MODE_DET)
    return w, mask, km


@pytest.mark.parametrize("bits", [1, 2, 3, 5, 7, 8, 9, 12])
def test_fixed_width_streams_match_the_device_layout(bits):
    rng = np.random.RandomState(bits)
    sym = rng.randint(0, 2 ** bits, size=1003).astype(np.uint32)
    packed = S.pack_fixed(sym, bits)
    assert packed.tobytes() == O.pack_codes(sym.astype(np.int32), bits).tobytes()  # same stream as nnc_kmeans1d_f32 emits
    assert np.array_equal(S.unpack_fixed(packed, sym.size, bits), sym)


def test_huffman_round_trip_and_prefix_property():
    rng = np.random.RandomState(0)
    for k, n in ((2, 50), (5, 1000), (16, 20000), (256, 50000), (300, 1000)):
        p = rng.dirichlet(np.ones(k) * 0.3)
        sym = rng.choice(k, size=n, p=p).astype(np.uint32)
        lengths = S.huffman_lengths(np.bincount(sym, minlength=k))
        used = lengths[lengths > 0].astype(np.float64)
        assert abs(np.sum(2.0 ** -used) - 1.0) < 1e-12 or used.size == 1  # Kraft equality: a complete prefix code
        payload = S.huffman_encode(sym, lengths)
        assert np.array_equal(S.huffman_decode(payload, n, lengths), sym)
        # never worse than the fixed-width stream by more than the rounding of the last byte, better on skewed data
        assert payload.size <= S.pack_fixed(sym, max(1, int(k - 1).bit_length())).size + 1
    one = np.zeros(77, dtype=np.uint32) + 3
    lengths = S.huffman_lengths(np.bincount(one, minlength=8))
    assert np.array_equal(S.huffman_decode(S.huffman_encode(one, lengths), 77, lengths), one)


@pytest.mark.parametrize("huffman", [False, True])
@pytest.mark.parametrize("layout", [S.DENSE, S.SPARSE])
def test_layer_round_trip_bit_exact(layout, huffman):
    w, mask, km = _lenet_layer()
    fit = _Fit(km.cluster_centers_, km.labels_, 4)
    layer = S.from_result(w.shape, fit, mask, layout=layout, rel_bits=5, huffman=huffman)
    buf = io.BytesIO()
    nbytes = S.save_compressed(buf, {"dense": layer})
    assert nbytes == buf.getbuffer().nbytes
    buf.seek(0)
    back = S.load_compressed(buf)["dense"]
    assert back.shape == w.shape and back.codebook.tobytes() == km.cluster_centers_.ravel().tobytes()
    ris = km.cluster_centers_.ravel()[km.labels_].reshape(w.shape)  # utility.py:239
    if layout == S.DENSE:
        assert back.dequantize().tobytes() == ris.tobytes()
        assert np.array_equal(back.pruning_mask(), mask)
    else:
        expect = np.where(mask, np.float32(0), ris)  # mask re-applied after quantisation (trainer.py:195-206)
        assert back.dequantize().tobytes() == expect.tobytes()
        assert np.array_equal(back.pruning_mask(), mask)
        assert int((back.gaps == 31).sum()) > 0  # the filler entries of figure 2 occur on this layer
    # the point of the format: far below 4 bytes per weight (dense: (4 + 1) / 32 of it; sparse + Huffman: less)
    assert nbytes < 0.2 * w.nbytes


def test_sparse_layout_edge_cases():
    cb = np.array([0.0, 1.5, -2.0], dtype=np.float32)
    for mask in (np.ones(40, bool), np.zeros(40, bool), np.r_[np.ones(39, bool), False], np.r_[False, np.ones(39, bool)]):
        codes = (np.arange(40) % 3).astype(np.uint32)
        gaps, ecodes = S.to_sparse(codes, mask, rel_bits=2)
        L = S.CompressedLayer((5, 8), cb, 2, S.SPARSE, codes=ecodes, gaps=gaps, rel_bits=2)
        buf = io.BytesIO()
        S.save_compressed(buf, {"x": L})
        buf.seek(0)
        back = S.load_compressed(buf)["x"]
        assert np.array_equal(back.pruning_mask().ravel(), mask)
        assert back.dequantize().ravel().tobytes() == np.where(mask, np.float32(0), cb[codes]).tobytes()


def test_bad_files_are_rejected():
    with pytest.raises(ValueError):
        S.load_compressed(io.BytesIO(b"XXXX\x01\x00\x00\x00"))
    with pytest.raises(ValueError):
        S.load_compressed(io.BytesIO(b"NNCL\x09\x00\x00\x00"))


@pytest.mark.gpu
def test_device_layer_to_file_and_back():
    import torch

    from neural_network_compression_b200.common import utility as U

    w = D.lenet5_tensors()[2][1].copy()
    t = torch.from_numpy(w.copy()).cuda()
    mask_bits, km = U.compress_weight(t, 1, True, 4, "linear", mask_bits=True)
    w_ref = w.copy()
    mask_ref = O.prune_weigth(w_ref, 1)
    assert mask_bits.dtype == torch.uint8 and mask_bits.numel() == (w.size + 7) // 8
    assert mask_bits.cpu().numpy().tobytes() == np.packbits(mask_ref.ravel(), bitorder="little").tobytes()
    # the stand-alone packer on a device bool mask and on a host array
    m2 = U.prune_weigth(torch.from_numpy(w.copy()).cuda(), 1)
    assert U.pack_mask_bits(m2).cpu().numpy().tobytes() == mask_bits.cpu().numpy().tobytes()
    assert U.pack_mask_bits(mask_ref).tobytes() == mask_bits.cpu().numpy().tobytes()
    odd = np.random.RandomState(1).rand(1003) < 0.4
    assert U.pack_mask_bits(odd).tobytes() == np.packbits(odd, bitorder="little").tobytes()
    for layout in (S.DENSE, S.SPARSE):
        layer = S.from_result(w.shape, km, mask_ref, layout=layout, huffman=True)
        buf = io.BytesIO()
        S.save_compressed(buf, {"dense": layer})
        buf.seek(0)
        back = S.load_compressed(buf)["dense"]
        deq = U.dequantize(km.packed_codes, w.size, km.code_bits, km.cluster_centers_).cpu().numpy().reshape(w.shape)
        expect = deq if layout == S.DENSE else np.where(mask_ref, np.float32(0), deq)
        assert back.dequantize().tobytes() == expect.tobytes()
