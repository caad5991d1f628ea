"""The file loaders against the vectors recorded from the reference's loaders (lddutils.py:150-229) and
the packed pass-through used with the fused GPU unpack."""
import io

import numpy as np

from lddecode_b200 import loaders, synth


def test_loaders_match_reference(golden):
    g = golden("unpack")
    f = io.BytesIO(g["r30_words"].tobytes())
    for s, exp in zip(g["r30_py_starts"], g["r30_py"]):
        assert np.array_equal(loaders.load_packed_data_3_32(f, int(s), 1000), exp)
    f = io.BytesIO(g["lds_bytes"].tobytes())
    for s, exp in zip(g["lds_py_starts"], g["lds_py"]):
        assert np.array_equal(loaders.load_packed_data_4_40(f, int(s), 1000), exp)
    # short reads -> None (lddutils.py:117-129)
    assert loaders.load_packed_data_4_40(f, len(g["lds_bytes"]) // 5 * 4 - 10, 1000) is None
    raw = np.arange(5000, dtype=np.uint8)
    f8 = io.BytesIO(raw.tobytes())
    assert np.array_equal(loaders.load_unpacked_data_u8(f8, 100, 50), raw[100:150])
    assert loaders.load_unpacked_data_u8(f8, 4990, 50) is None
    s16 = (np.arange(3000) - 1500).astype('<i2')
    f16 = io.BytesIO(s16.tobytes())
    assert np.array_equal(loaders.load_unpacked_data_s16(f16, 7, 100), s16[7:107])


def test_raw_packed_ranges_cover_the_request():
    rng = np.random.default_rng(0)
    s10 = rng.integers(0, 1024, 12 * 400, dtype=np.uint16)
    fr = io.BytesIO(synth.pack_r30(s10).tobytes())
    fl = io.BytesIO(synth.pack_lds(s10).tobytes())
    from oracle import ldd_oracle as O
    for sample, n in ((0, 100), (1, 100), (2, 999), (1234, 1000)):
        w, first = loaders.raw_r30(fr, sample, n)
        assert first % 3 == 0 and first <= sample
        assert np.array_equal(O.unpack_r30_raw(w, sample - first, n), s10[sample:sample + n].astype(np.int16))
        b, first = loaders.raw_lds(fl, sample, n)
        assert first % 4 == 0 and first <= sample
        assert np.array_equal(O.unpack_lds(b, sample - first, n), s10[sample:sample + n])
    assert loaders.raw_r30(fr, len(s10) - 10, 100)[0] is None
