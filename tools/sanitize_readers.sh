#!/bin/sh
# Build the host reader driver (tReaders) with AddressSanitizer + UBSan, and again with
# ThreadSanitizer, and run the CPU reader tests / threaded decode paths against them.
# Round-1 record: 53 reader tests pass under ASan+UBSan (abort_on_error), TSan reports no race
# in the plane- / chunk-parallel decode with OI_IO_THREADS=6.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
cd "$ROOT/openimpala_b200/host"
SRCS="apps/tReaders.cpp props/TortuosityHypre.cpp props/EffectiveDiffusivityHypre.cpp props/VolumeFraction.cpp io/TiffReader.cpp io/RawReader.cpp io/HDF5Reader.cpp io/DatReader.cpp"
LINK="-L../lib -lopenimpala_b200 -lz -pthread -Wl,-rpath,$ROOT/openimpala_b200/lib"
g++ -O1 -g -std=c++17 -fsanitize=address,undefined -fno-omit-frame-pointer -Iamrex_shim -I../../include -o /tmp/tReaders_asan $SRCS $LINK
g++ -O1 -g -std=c++17 -fsanitize=thread -Iamrex_shim -I../../include -o /tmp/tReaders_tsan $SRCS $LINK
cd "$ROOT"
cp openimpala_b200/bin/tReaders /tmp/tReaders_orig
trap 'cp /tmp/tReaders_orig openimpala_b200/bin/tReaders' EXIT
cp /tmp/tReaders_asan openimpala_b200/bin/tReaders
ASAN_OPTIONS=protect_shadow_gap=0:detect_leaks=0:abort_on_error=1 UBSAN_OPTIONS=halt_on_error=1 \
    python -m pytest tests/test_host_apps.py -x -q -m "not gpu"
for args in "mode=tiff tifffile=tests/golden/SampleData_2Phase_stack_3d_1bit.tif u8_chunk=7" \
            "mode=hdf5 hdf5file=tests/golden/SampleData_2Phase_3d.hdf5 hdf5dataset=image u8_chunk=11" \
            "mode=raw rawfile=tests/golden/SampleData_2Phase_stack_3d_uint8.raw width=100 height=100 depth=100 datatype=UINT8 u8_chunk=9"; do
    OI_IO_THREADS=6 /tmp/tReaders_tsan $args gpu_count=0 2>&1 | grep -E "WARNING|TEST|U8Chunk"
done
