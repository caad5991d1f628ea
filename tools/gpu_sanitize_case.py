"""Small end-to-end case for compute-sanitizer: one NTSC and one PAL field through every kernel
(demod in all three lanes, peaks, hsync, burst/pilot, TBC, audio phase 2, unpackers)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from lddecode_b200 import _lib, field as F, pipeline, rfdecode, synth

for name in ("ntsc", "pal"):
    g = dict(np.load(os.path.join(ROOT, "tests/golden/%s.npz" % name)))
    system = "PAL" if name == "pal" else "NTSC"
    cap = g["capture"]
    for prec in ("f64", "f32", "mixed"):
        rf = rfdecode.RFDecode(float(g["fs_mhz"]), system, int(g["blocklen"]), precision=prec)
        be = rf._be
        dd = rf.demod_device(be.to_device(cap), _lib.FMT_U8, 0, len(cap), 0, int(g["demod_length"]), 1)
        f = (F.FieldNTSC if system == "NTSC" else F.FieldPAL)(rf, dd, 0)
        d = f.dspicture.astype(int) - g["field_dspicture"].astype(int)
        print(name, prec, "valid", f.valid, "peaks ok", np.array_equal(f.peaklist, g["field_peaklist"]), "tbc maxdiff", np.abs(d).max(), flush=True)
    cd = pipeline.CaptureDecoder(rf, max_fields=16)
    res = cd.decode(be.to_device(cap), _lib.FMT_U8, len(cap))
    print(name, "pipeline fields", len(cd.pictures(res)))
u = dict(np.load(os.path.join(ROOT, "tests/golden/unpack.npz")))
lib = _lib.load()
w = torch.from_numpy(u["r30_words"].astype(np.int32)).cuda()
o = torch.empty(len(u["r30_words"]) * 3, dtype=torch.int16, device="cuda")
lib.ldd_unpack_r30_ddunpack(w.data_ptr(), len(u["r30_words"]), o.data_ptr(), None)
torch.cuda.synchronize()
print("unpack ok", np.array_equal(o.cpu().numpy(), u["r30_ddunpack_i16"]))
