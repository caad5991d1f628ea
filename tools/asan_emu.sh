#!/bin/sh
# Builds the CPU emulation of the kernel sources with AddressSanitizer and runs the NTSC + PAL field
# chain through it (compute-sanitizer is closed on the GPU pool; this is the memory-safety check we can run).
set -e
cd "$(dirname "$0")/.."
mkdir -p /tmp/ldd_asan
for f in lddecode_b200/csrc/*.cu; do
  g++ -std=c++17 -O1 -g -fPIC -fsanitize=address -fno-omit-frame-pointer -DLDD_EMU -I tests/emu -I lddecode_b200/csrc \
      -x c++ -c "$f" -o /tmp/ldd_asan/$(basename "$f").o -Wno-unknown-pragmas &
done
g++ -std=c++17 -O1 -g -fPIC -fsanitize=address -c tests/emu/cuda_emu.cpp -o /tmp/ldd_asan/emu.o
wait
g++ -shared -fsanitize=address -o /tmp/ldd_asan/libldd_emu_asan.so /tmp/ldd_asan/*.o
LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 \
python - <<'PY'
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from lddecode_b200._backend import EmuBackend
from lddecode_b200 import rfdecode, field, pipeline, _lib
be = EmuBackend('/tmp/ldd_asan/libldd_emu_asan.so')
for name in ('ntsc', 'pal'):
    g = dict(np.load('tests/golden/%s.npz' % name))
    system = 'PAL' if name == 'pal' else 'NTSC'
    for prec in ('f64', 'mixed'):
        rf = rfdecode.RFDecode(float(g['fs_mhz']), system, int(g['blocklen']), precision=prec, _backend=be)
        cap = g['capture']
        dd = rf.demod_device(be.to_device(cap), 0, 0, len(cap), 0, int(g['demod_length']), 1)
        f = (field.FieldNTSC if system == 'NTSC' else field.FieldPAL)(rf, dd, 0)
        print(name, prec, 'valid', f.valid)
    res = pipeline.CaptureDecoder(rf, max_fields=16).decode(be.to_device(cap), 0, len(cap))
    print(name, 'pipeline windows', res.nwindows)
# the reference's default rate from packed .lds bytes, both audio channels, range PCM (longest lines, other staging windows)
from lddecode_b200 import synth
for system in ('NTSC', 'PAL'):
    n = int(40e6 / (30 if system == 'NTSC' else 25) * 1.1) // 4 * 4
    s10 = synth.SynthRF(system, 40.0, seed=5, bits=10).generate(n)
    rf = rfdecode.RFDecode(40.0, system, 16384, _backend=be)
    cd = pipeline.CaptureDecoder(rf, max_fields=16)
    res = cd.decode(be.to_device(synth.pack_lds(s10)), _lib.FMT_LDS40, n)
    pcm, _, _ = cd.pcm(res, chain='fields')
    print(system, '40 MSPS .lds: located', len(res.located), 'pcm', [None if p is None else len(p) for p in pcm])
print('ASAN_CLEAN')
PY
