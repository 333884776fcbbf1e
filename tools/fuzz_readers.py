"""Mutation fuzz of the host readers under AddressSanitizer + UBSan: truncations and random
byte flips of TIFF (uncompressed, LZW) and HDF5 (contiguous, chunked + deflate / shuffle /
fletcher32) files through /tmp/tReaders_asan (built by tools/sanitize_readers.sh).  A clean
rejection (non-zero exit, amrex::Abort) is fine; a sanitizer report, a signal or a hang is an
issue.  Round-1 record: FUZZ_SEED=5 (300 inputs) and FUZZ_SEED=11 FUZZ_TRIALS=240 (1200 inputs):
no memory error and no hang after the directory-cycle / extent checks went in; the only
non-clean exits left are out-of-memory aborts on headers that claim absurd extents.
    sh tools/sanitize_readers.sh && FUZZ_SEED=11 FUZZ_TRIALS=240 python tools/fuzz_readers.py"""
import os
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import os, random, subprocess, sys, shutil
sys.path.insert(0, os.path.join(_ROOT, 'tests')); sys.path.insert(0, _ROOT)
import numpy as np
import test_host_apps as t
random.seed(int(os.environ.get("FUZZ_SEED", "5")))
rng = np.random.default_rng(3)
os.makedirs('/tmp/fuzz', exist_ok=True)
vol = rng.integers(0, 255, (11, 13, 17)).astype('u1')
t._write_hdf5('/tmp/fuzz/base.h5', 'image', vol, chunks=(4, 5, 6), gzip=True, shuffle=False)
vol16 = rng.integers(0, 60000, (11, 13, 17)).astype('<u2')
t._write_hdf5('/tmp/fuzz/base16.h5', 'image', vol16, chunks=(4, 5, 6), gzip=True, shuffle=True, fletcher=True)
bases = [('/tmp/fuzz/base.h5', ['mode=hdf5', 'hdf5dataset=image']), ('/tmp/fuzz/base16.h5', ['mode=hdf5', 'hdf5dataset=image']),
         (os.path.join(_ROOT, 'tests/golden/SampleData_2Phase_3d.hdf5'), ['mode=hdf5', 'hdf5dataset=image']),
         (os.path.join(_ROOT, 'tests/golden/SampleData_2Phase_squared.tif'), ['mode=tiff']),
         ('/tmp/pack64_lzw.tif', ['mode=tiff'])]
from PIL import Image
ims=[Image.fromarray(rng.integers(0,255,(64,64)).astype(np.uint8)) for _ in range(8)]
ims[0].save('/tmp/pack64_lzw.tif', save_all=True, append_images=ims[1:], compression='tiff_lzw')
env = dict(os.environ, ASAN_OPTIONS='protect_shadow_gap=0:detect_leaks=0:abort_on_error=0:exitcode=99', UBSAN_OPTIONS='halt_on_error=0')
bad = 0; runs = 0
for base, args in bases:
    data = open(base, 'rb').read()
    for trial in range(int(os.environ.get("FUZZ_TRIALS", "60"))):
        b = bytearray(data)
        kind = trial % 3
        if kind == 0:
            b = b[:random.randrange(8, len(b))]
        elif kind == 1:
            for _ in range(random.randrange(1, 6)):
                b[random.randrange(0, min(len(b), 4096))] = random.randrange(256)
        else:
            for _ in range(random.randrange(1, 10)):
                b[random.randrange(0, len(b))] = random.randrange(256)
        p = '/tmp/fuzz/f' + os.path.splitext(base)[1]
        open(p, 'wb').write(b)
        key = 'hdf5file' if 'hdf5' in args[0] else 'tifffile'
        try:
            r = subprocess.run(['/tmp/tReaders_asan', *args, f'{key}={p}', 'gpu_count=0', 'u8_chunk=3'], capture_output=True, text=True, env=env, timeout=60)
        except subprocess.TimeoutExpired:
            print('TIMEOUT', base, trial); bad += 1; continue
        runs += 1
        if 'AddressSanitizer' in r.stderr or 'runtime error' in r.stderr or r.returncode in (99, -11, -6, -8):
            bad += 1
            print('ISSUE', base, trial, kind, r.returncode, r.stderr[-600:])
            shutil.copy(p, f'/tmp/fuzz/crash_{os.path.basename(base)}_{trial}')
print('runs', runs, 'issues', bad)
