"""On-disk / wire format of a compressed layer.

The reference keeps nothing: masks live in a class-level dict (neural_network_compression/common/trainer.py:25) and
the quantised weights are written back into the Keras layer as dense float32 (trainer.py:70), so the 2..8-bit
codebook structure is lost as soon as `Trainer.quantize` returns.  This module serialises what the hot path produces
-- codebook (k x float32), n-bit cluster indices, pruning mask -- in the two layouts of Han et al., "Deep Compression"
(papers/deep-compression.pdf, section 3 and figure 2), optionally Huffman coded (section 4):

    DENSE    one code per weight, `code_bits` each, little-endian bit stream (what `nnc_kmeans1d_f32` / `nnc_compress_f32`
             emit), plus the 1-bit pruning mask.  Decodes to exactly the reference's `ris = cluster_centers_[labels_]`
             (utility.py:239): pruned weights come back as the centroid that captured the zeros, as in the reference.
    SPARSE   relative-indexed non-zeros: for every surviving weight the gap to the previous survivor in `rel_bits` bits and
             its code; a gap that does not fit emits a filler entry (gap = 2^rel_bits - 1, code of the filler is ignored)
             exactly like the padding zero of the paper's figure 2.  Decodes to the pruned-and-quantised tensor with exact
             zeros at the pruned positions (mask re-applied: trainer.py:195-206 after quantisation).

File layout (little endian):
    magic "NNCL" | u16 version = 1 | u16 n_layers
    per layer:  u16 name_len | name utf-8 | u8 layout | u8 code_bits | u8 rel_bits | u8 huffman | u32 k | u8 ndim | u64 shape[ndim]
                | f32 codebook[k] | sections...
      DENSE  : [codes] [mask bits: ceil(n / 8) bytes, bit i of byte i / 8 = mask of weight i]  (mask optional: u8 has_mask)
      SPARSE : u64 n_entries | [gaps] [codes]
    a [stream] is: u64 n_symbols | u64 n_bytes | (huffman: u16 n_lengths | u8 code_length[n_lengths]) | payload
Everything is integer / byte work; the round trip is bit exact (tests/test_storage.py).

The bit packing of the mask runs on the device (`nnc_pack_bits_u8`) when the mask lives there, so a compressed layer
leaves the GPU as n * (code_bits + 1) / 8 bytes instead of n * 5.
"""
from __future__ import annotations

import heapq
import io
import struct
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np

MAGIC = b"NNCL"
VERSION = 1
DENSE, SPARSE = 0, 1
MAX_HUFFMAN_LENGTH = 24


# ---------------------------------------------------------------------------------------------------------
# fixed-width bit streams (the layout of the device's packed codes: weight i occupies bits [i * b, (i + 1) * b))
# ---------------------------------------------------------------------------------------------------------
def pack_fixed(symbols: np.ndarray, bits: int) -> np.ndarray:
    symbols = np.ascontiguousarray(symbols, dtype=np.uint32)
    if bits == 8:
        return symbols.astype(np.uint8)
    n = symbols.size
    shifts = np.arange(bits, dtype=np.uint32)
    bitmat = ((symbols[:, None] >> shifts[None, :]) & 1).astype(np.uint8)  # little endian inside a symbol
    return np.packbits(bitmat.reshape(-1), bitorder="little")[: (n * bits + 7) // 8]


def unpack_fixed(payload: np.ndarray, n: int, bits: int) -> np.ndarray:
    payload = np.ascontiguousarray(payload, dtype=np.uint8)
    if bits == 8:
        return payload[:n].astype(np.uint32)
    bitvec = np.unpackbits(payload, bitorder="little")[: n * bits].reshape(n, bits).astype(np.uint32)
    return (bitvec << np.arange(bits, dtype=np.uint32)[None, :]).sum(axis=1, dtype=np.uint32)


# ---------------------------------------------------------------------------------------------------------
# canonical Huffman coding (Deep Compression section 4); code lengths travel, codes are rebuilt on both sides
# ---------------------------------------------------------------------------------------------------------
def huffman_lengths(hist: np.ndarray) -> np.ndarray:
    """Code length per symbol (0 for symbols that do not occur).  Lengths above MAX_HUFFMAN_LENGTH are avoided by
    flattening the histogram (halving counts, keeping them >= 1) until the tree is shallow enough."""
    hist = np.asarray(hist, dtype=np.int64).copy()
    used = np.flatnonzero(hist > 0)
    lengths = np.zeros(hist.size, dtype=np.uint8)
    if used.size == 0:
        return lengths
    if used.size == 1:
        lengths[used[0]] = 1
        return lengths
    while True:
        heap = [(int(hist[s]), int(s), (int(s),)) for s in used]
        heapq.heapify(heap)
        depth = dict.fromkeys((int(s) for s in used), 0)
        while len(heap) > 1:
            c1, t1, m1 = heapq.heappop(heap)
            c2, t2, m2 = heapq.heappop(heap)
            for s in m1 + m2:
                depth[s] += 1
            heapq.heappush(heap, (c1 + c2, min(t1, t2), m1 + m2))
        if max(depth.values()) <= MAX_HUFFMAN_LENGTH:
            break
        hist[used] = np.maximum(1, hist[used] // 2)
    for s, d in depth.items():
        lengths[s] = d
    return lengths


def canonical_codes(lengths: np.ndarray):
    """(code, length) per symbol: canonical assignment, symbols ordered by (length, symbol); codes are stored MSB first."""
    lengths = np.asarray(lengths, dtype=np.int64)
    codes = np.zeros(lengths.size, dtype=np.uint64)
    code = 0
    prev_len = 0
    for s in sorted(np.flatnonzero(lengths > 0), key=lambda i: (lengths[i], i)):
        code <<= int(lengths[s] - prev_len)
        codes[s] = code
        code += 1
        prev_len = int(lengths[s])
    return codes, lengths


def huffman_encode(symbols: np.ndarray, lengths: np.ndarray) -> np.ndarray:
    codes, lengths = canonical_codes(lengths)
    symbols = np.ascontiguousarray(symbols, dtype=np.int64)
    ln = lengths[symbols]
    if symbols.size and int(ln.min()) == 0:
        raise ValueError("symbol without a Huffman code")
    total = int(ln.sum())
    ends = np.cumsum(ln)
    starts = ends - ln
    out_bits = np.zeros(total, dtype=np.uint8)
    cw = codes[symbols]
    for b in range(int(ln.max()) if symbols.size else 0):  # bit b (from the MSB) of every code word long enough
        sel = ln > b
        out_bits[starts[sel] + b] = ((cw[sel] >> (ln[sel] - 1 - b).astype(np.uint64)) & np.uint64(1)).astype(np.uint8)
    return np.packbits(out_bits, bitorder="big")


def huffman_decode(payload: np.ndarray, n: int, lengths: np.ndarray) -> np.ndarray:
    codes, lengths = canonical_codes(lengths)
    used = np.flatnonzero(lengths > 0)
    out = np.empty(n, dtype=np.uint32)
    if n == 0:
        return out
    maxlen = int(lengths[used].max())
    # window table: the symbol and length of the code word that prefixes every maxlen-bit window
    table_sym = np.zeros(1 << maxlen, dtype=np.uint32)
    table_len = np.zeros(1 << maxlen, dtype=np.uint8)
    for s in used:
        l = int(lengths[s])
        lo = int(codes[s]) << (maxlen - l)
        table_sym[lo:lo + (1 << (maxlen - l))] = s
        table_len[lo:lo + (1 << (maxlen - l))] = l
    bits = np.unpackbits(np.ascontiguousarray(payload, dtype=np.uint8), bitorder="big")
    bits = np.concatenate([bits, np.zeros(maxlen, dtype=np.uint8)])
    # value of the maxlen-bit window starting at every bit position
    win = np.zeros(bits.size - maxlen + 1, dtype=np.uint32)
    for b in range(maxlen):
        win = (win << np.uint32(1)) | bits[b:b + win.size]
    sym_at = table_sym[win].tolist()
    len_at = table_len[win].tolist()
    p = 0
    res = [0] * n
    for i in range(n):  # the chain of code-word starts is sequential by nature
        res[i] = sym_at[p]
        p += len_at[p]
    out[:] = res
    return out


# ---------------------------------------------------------------------------------------------------------
# layer container
# ---------------------------------------------------------------------------------------------------------
@dataclass
class CompressedLayer:
    shape: tuple
    codebook: np.ndarray  # (k,) float32
    code_bits: int
    layout: int = DENSE
    codes: Optional[np.ndarray] = None  # DENSE: uint32 code per weight; SPARSE: code per entry (fillers included)
    mask: Optional[np.ndarray] = None  # DENSE: bool per weight (True = pruned), optional
    gaps: Optional[np.ndarray] = None  # SPARSE: uint32 gap per entry
    rel_bits: int = 0
    huffman: bool = False

    @property
    def n(self) -> int:
        return int(np.prod(self.shape)) if len(self.shape) else 1

    def dequantize(self) -> np.ndarray:
        """DENSE: cluster_centers_[labels_] (utility.py:239).  SPARSE: the same with exact zeros at the pruned positions."""
        if self.layout == DENSE:
            return self.codebook[self.codes].reshape(self.shape)
        out = np.zeros(self.n, dtype=np.float32)
        filler = self.gaps == (1 << self.rel_bits) - 1
        pos = np.cumsum(self.gaps.astype(np.int64) + np.where(filler, 0, 1)) - np.where(filler, 0, 1)
        real = ~filler
        out[pos[real]] = self.codebook[self.codes[real]]
        return out.reshape(self.shape)

    def pruning_mask(self) -> Optional[np.ndarray]:
        if self.layout == DENSE:
            return None if self.mask is None else self.mask.reshape(self.shape)
        m = np.ones(self.n, dtype=bool)
        filler = self.gaps == (1 << self.rel_bits) - 1
        pos = np.cumsum(self.gaps.astype(np.int64) + np.where(filler, 0, 1)) - np.where(filler, 0, 1)
        m[pos[~filler]] = False
        return m.reshape(self.shape)

    def payload_bytes(self) -> int:
        buf = io.BytesIO()
        _write_layer(buf, "", self)
        return buf.getbuffer().nbytes


def to_sparse(codes: np.ndarray, mask: np.ndarray, rel_bits: int):
    """(gaps, entry codes) of the relative-indexed layout: survivors = ~mask in flattened order; a gap of
    2^rel_bits - 1 or more emits filler entries (paper figure 2)."""
    mask = np.asarray(mask, dtype=bool).reshape(-1)
    codes = np.asarray(codes, dtype=np.uint32).reshape(-1)
    pos = np.flatnonzero(~mask)
    prev = np.concatenate([[-1], pos[:-1]])
    gap = pos - prev - 1
    cap = (1 << rel_bits) - 1
    n_fill = gap // cap  # every filler advances by `cap` positions
    total = int(pos.size + n_fill.sum())
    gaps = np.full(total, cap, dtype=np.uint32)
    ecodes = np.zeros(total, dtype=np.uint32)
    at = np.cumsum(n_fill + 1) - 1  # entry index of every real survivor
    gaps[at] = (gap - n_fill * cap).astype(np.uint32)
    ecodes[at] = codes[pos]
    return gaps, ecodes


def from_result(shape, kmeans, mask=None, layout: int = DENSE, rel_bits: int = 5, huffman: bool = False) -> CompressedLayer:
    """Build a CompressedLayer from what `utility.compress_weight` / `get_quantized_weight` return (`KMeansResult` with
    packed_codes; host arrays or device tensors)."""
    n = int(np.prod(shape)) if len(shape) else 1
    packed = kmeans.packed_codes
    if hasattr(packed, "detach"):
        packed = packed.detach().cpu().numpy()
    codes = unpack_fixed(np.asarray(packed), n, kmeans.code_bits)
    m = None
    if mask is not None:
        m = mask.detach().cpu().numpy() if hasattr(mask, "detach") else np.asarray(mask)
        m = m.astype(bool).reshape(-1)
    codebook = np.ascontiguousarray(np.asarray(kmeans.cluster_centers_, dtype=np.float32).ravel())
    if layout == DENSE:
        return CompressedLayer(tuple(shape), codebook, kmeans.code_bits, DENSE, codes=codes, mask=m, huffman=huffman)
    if m is None:
        raise ValueError("the sparse layout needs the pruning mask")
    gaps, ecodes = to_sparse(codes, m, rel_bits)
    return CompressedLayer(tuple(shape), codebook, kmeans.code_bits, SPARSE, codes=ecodes, gaps=gaps, rel_bits=rel_bits, huffman=huffman)


# ---------------------------------------------------------------------------------------------------------
# serialisation
# ---------------------------------------------------------------------------------------------------------
def _write_stream(f, symbols: np.ndarray, bits: int, huffman: bool):
    symbols = np.ascontiguousarray(symbols, dtype=np.uint32)
    if huffman:
        lengths = huffman_lengths(np.bincount(symbols, minlength=1 << bits))
        payload = huffman_encode(symbols, lengths)
        f.write(struct.pack("<QQH", symbols.size, payload.size, lengths.size))
        f.write(lengths.tobytes())
    else:
        payload = pack_fixed(symbols, bits)
        f.write(struct.pack("<QQ", symbols.size, payload.size))
    f.write(payload.tobytes())


def _read_stream(f, bits: int, huffman: bool) -> np.ndarray:
    n, nbytes = struct.unpack("<QQ", f.read(16))
    if huffman:
        (nl,) = struct.unpack("<H", f.read(2))
        lengths = np.frombuffer(f.read(nl), dtype=np.uint8)
        payload = np.frombuffer(f.read(nbytes), dtype=np.uint8)
        return huffman_decode(payload, n, lengths)
    payload = np.frombuffer(f.read(nbytes), dtype=np.uint8)
    return unpack_fixed(payload, n, bits)


def _write_layer(f, name: str, L: CompressedLayer):
    nm = name.encode("utf-8")
    f.write(struct.pack("<H", len(nm)))
    f.write(nm)
    f.write(struct.pack("<BBBBIB", L.layout, L.code_bits, L.rel_bits, int(L.huffman), L.codebook.size, len(L.shape)))
    f.write(struct.pack("<%dQ" % len(L.shape), *L.shape))
    f.write(np.ascontiguousarray(L.codebook, dtype="<f4").tobytes())
    if L.layout == DENSE:
        _write_stream(f, L.codes, L.code_bits, L.huffman)
        f.write(struct.pack("<B", 0 if L.mask is None else 1))
        if L.mask is not None:
            f.write(np.packbits(L.mask.reshape(-1).astype(np.uint8), bitorder="little").tobytes())
    else:
        f.write(struct.pack("<Q", L.codes.size))
        _write_stream(f, L.gaps, L.rel_bits, L.huffman)
        _write_stream(f, L.codes, L.code_bits, L.huffman)


def _read_layer(f):
    (nl,) = struct.unpack("<H", f.read(2))
    name = f.read(nl).decode("utf-8")
    layout, code_bits, rel_bits, huff, k, ndim = struct.unpack("<BBBBIB", f.read(9))
    shape = struct.unpack("<%dQ" % ndim, f.read(8 * ndim))
    codebook = np.frombuffer(f.read(4 * k), dtype="<f4").astype(np.float32)
    L = CompressedLayer(tuple(int(s) for s in shape), codebook, code_bits, layout, rel_bits=rel_bits, huffman=bool(huff))
    if layout == DENSE:
        L.codes = _read_stream(f, code_bits, bool(huff))
        (has_mask,) = struct.unpack("<B", f.read(1))
        if has_mask:
            L.mask = np.unpackbits(np.frombuffer(f.read((L.n + 7) // 8), dtype=np.uint8), bitorder="little")[: L.n].astype(bool)
    elif layout == SPARSE:
        (n_entries,) = struct.unpack("<Q", f.read(8))
        L.gaps = _read_stream(f, rel_bits, bool(huff))
        L.codes = _read_stream(f, code_bits, bool(huff))
        if L.gaps.size != n_entries or L.codes.size != n_entries:
            raise ValueError("corrupt sparse layer %r" % name)
    else:
        raise ValueError("unknown layout %d" % layout)
    return name, L


def save_compressed(file, layers: Dict[str, CompressedLayer]) -> int:
    """Write {name: CompressedLayer} to a path or a binary file object; returns the number of bytes written."""
    own = isinstance(file, (str, bytes))
    f = open(file, "wb") if own else file
    try:
        start = f.tell()
        f.write(MAGIC + struct.pack("<HH", VERSION, len(layers)))
        for name, L in layers.items():
            _write_layer(f, name, L)
        return f.tell() - start
    finally:
        if own:
            f.close()


def load_compressed(file) -> Dict[str, CompressedLayer]:
    own = isinstance(file, (str, bytes))
    f = open(file, "rb") if own else file
    try:
        if f.read(4) != MAGIC:
            raise ValueError("not a compressed-layer file (bad magic)")
        version, n_layers = struct.unpack("<HH", f.read(4))
        if version != VERSION:
            raise ValueError("unsupported version %d" % version)
        return dict(_read_layer(f) for _ in range(n_layers))
    finally:
        if own:
            f.close()
