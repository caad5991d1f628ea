"""File loaders with the reference's contract (lddutils.py:117-229): loader(infile, sample, readlen)
-> array of `readlen` samples starting at `sample`, or None when the file is too short.

Two flavours per format:
  load_*          host numpy, same values as the reference's loaders (they feed RFDecode.demod through
                  the module-global `rfdecode.loader`, exactly like lddecode.py:53-58 assigns them);
  raw_*           return the PACKED bytes of the range plus the sample offset, for RFDecode.demod_device
                  with LDD_FMT_R30 / LDD_FMT_LDS40: the 10-bit unpack then happens on the GPU inside the
                  block load of the demodulation kernel and the host never touches the samples.
"""
import numpy as np


def _read(infile, start, nbytes):
    infile.seek(start, 0)
    buf = infile.read(nbytes)
    return buf if len(buf) == nbytes else None


def load_unpacked_data(infile, sample, readlen, sampletype):
    """lddutils.py:131-141 (sampletype 1: uint8 cxadc data, 2: int16)."""
    buf = _read(infile, sample * sampletype, readlen * sampletype)
    if buf is None:
        return None
    return np.frombuffer(buf, 'int16' if sampletype == 2 else 'uint8')


def load_unpacked_data_u8(infile, sample, readlen):
    return load_unpacked_data(infile, sample, readlen, 1)


def load_unpacked_data_s16(infile, sample, readlen):
    return load_unpacked_data(infile, sample, readlen, 2)


def load_packed_data_3_32(infile, sample, readlen):
    """lddutils.py:150-173: 3 x 10 bit per little-endian u32 (.r30), raw 0..1023 as int16."""
    start = (sample // 3) * 4
    offset = sample % 3
    needed = int(np.ceil(readlen * 3 / 4) * 4) + 4
    infile.seek(start)
    buf = infile.read(needed)
    w = np.frombuffer(buf[:len(buf) // 4 * 4], '<u4')
    out = np.empty(len(w) * 3, dtype=np.int16)
    out[0::3] = w & 0x3ff
    out[1::3] = (w >> 10) & 0x3ff
    out[2::3] = (w >> 20) & 0x3ff
    res = out[offset:offset + readlen]
    return res if len(res) == readlen else None


def load_packed_data_4_40(infile, sample, readlen):
    """lddutils.py:195-229: 4 x 10 bit in 5 bytes, MSB first (.lds), raw 0..1023 as uint16."""
    start = (sample // 4) * 5
    offset = sample % 4
    needed = int(np.ceil(readlen * 5 // 4)) + 5
    infile.seek(start)
    buf = infile.read(needed)
    b = np.frombuffer(buf[:len(buf) // 5 * 5], 'uint8').astype(np.uint16).reshape(-1, 5)
    out = np.empty((len(b), 4), dtype=np.uint16)
    out[:, 0] = (b[:, 0] << 2) | (b[:, 1] >> 6)
    out[:, 1] = ((b[:, 1] & 0x3f) << 4) | (b[:, 2] >> 4)
    out[:, 2] = ((b[:, 2] & 0x0f) << 6) | (b[:, 3] >> 2)
    out[:, 3] = ((b[:, 3] & 0x03) << 8) | b[:, 4]
    res = out.reshape(-1)[offset:offset + readlen]
    return res if len(res) == readlen else None


def raw_r30(infile, sample, readlen):
    """Packed .r30 words covering [sample, sample+readlen): (uint32 array, first sample of the array)."""
    w0 = sample // 3
    nw = (sample + readlen + 2) // 3 - w0
    buf = _read(infile, w0 * 4, nw * 4)
    return (None, 0) if buf is None else (np.frombuffer(buf, '<u4'), w0 * 3)


def raw_lds(infile, sample, readlen):
    """Packed .lds bytes covering [sample, sample+readlen): (uint8 array, first sample of the array)."""
    g0 = sample // 4
    ng = (sample + readlen + 3) // 4 - g0
    buf = _read(infile, g0 * 5, ng * 5)
    return (None, 0) if buf is None else (np.frombuffer(buf, 'uint8'), g0 * 4)
