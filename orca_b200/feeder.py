"""
Packed-base feeder: 1 byte per bp instead of the reference's (L, 4) float32 one-hot.

The reference turns sequence text into a float32 one-hot on the host (selene_utils2.py:125-128 builds the whole
genome as a (4, N) float32 memmap through selene_sdk's `sequence_to_encoding`; `get_encoding_from_coords`
:216-230 slices / pads it) and `genomepredict` uploads that array once per strand and model, after materialising
the reverse complement with `sequence[:, ::-1, ::-1].copy()` (orca_predict.py:324-337): 16 B/bp over PCIe each time.

Here the sequence travels as bytes -- either codes 0..4 (A, C, G, T, N) or the raw ASCII of a FASTA record -- and
the first encoder kernel expands them on the fly (csrc/seq_in.cuh); the reverse-complement strand is read from the
same device buffer.  Mapping (identical to the reference feeder's): A/a, C/c, G/g, T/t -> one-hot in ACGT order,
anything else (N, IUPAC codes, pads) -> 0.25 in all four channels.

`to_onehot` is the host restatement of that mapping (used by tests and as documentation); `from_onehot` packs an
existing reference-format array so callers holding one can still upload 1 B/bp.
"""
import numpy as np

CODE_A, CODE_C, CODE_G, CODE_T, CODE_N = 0, 1, 2, 3, 4

_LUT = np.full(256, CODE_N, dtype=np.uint8)
for _i in range(5):
    _LUT[_i] = _i
for _ch, _code in (("A", CODE_A), ("C", CODE_C), ("G", CODE_G), ("T", CODE_T)):
    _LUT[ord(_ch)] = _code
    _LUT[ord(_ch.lower())] = _code

_ONEHOT = np.concatenate([np.eye(4, dtype=np.float32), np.full((1, 4), 0.25, dtype=np.float32)], axis=0)


def as_bases(sequence):
    """str / bytes / uint8 array -> uint8 array viewing the same bytes (no per-base work on the host).

    Text is passed through as ASCII: the device kernels understand both ASCII and codes."""
    if isinstance(sequence, str):
        sequence = sequence.encode("ascii")
    if isinstance(sequence, (bytes, bytearray, memoryview)):
        return np.frombuffer(sequence, dtype=np.uint8)
    arr = np.asarray(sequence)
    if arr.dtype != np.uint8:
        raise TypeError("packed sequences must be uint8 (codes 0..4 or ASCII), got %s" % arr.dtype)
    return arr


def codes(sequence):
    """Normalise to codes 0..4 (host LUT); the device does the same mapping itself, this is for host-side use."""
    return _LUT[as_bases(sequence)]


def to_onehot(sequence):
    """Packed bases (..., L) -> float32 one-hot (..., L, 4) exactly as the reference feeder encodes them."""
    return _ONEHOT[codes(sequence)]


def from_onehot(onehot, strict=True):
    """(..., L, 4) reference-format array -> uint8 codes (..., L).

    Rows must be a one-hot base or the 0.25 'unknown' row; with strict=True anything else raises ValueError
    (such inputs have no packed form and must take the fp32 path)."""
    a = np.asarray(onehot)
    if a.shape[-1] != 4:
        raise ValueError("expected (..., L, 4), got %s" % (a.shape,))
    idx = a.argmax(axis=-1).astype(np.uint8)
    top = np.take_along_axis(a, idx[..., None].astype(np.int64), axis=-1)[..., 0]
    is_base = (top == 1) & (a.sum(axis=-1) == 1)
    is_n = np.all(a == 0.25, axis=-1)
    if strict and not np.all(is_base | is_n):
        raise ValueError("array has rows that are neither one-hot bases nor 0.25 'N' rows; no packed form")
    return np.where(is_base, idx, np.uint8(CODE_N)).astype(np.uint8)


def reverse_complement(sequence):
    """Host restatement of the strand flip on packed codes (A<->T, C<->G, N stays); the device path does not
    materialise this -- it walks the buffer backwards."""
    c = codes(sequence)[..., ::-1]
    return np.where(c < 4, 3 - c, c).astype(np.uint8)
