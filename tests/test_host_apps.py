"""The C++ host front end (openimpala_b200/host): reference-named classes over
AMReX-shaped shims, readers, and the Diffusion / tTortuosity drivers.  CPU tests
cover the readers (reference sample files: tTiffReader / tRawReader /
tHDF5Reader facts); GPU tests run the apps end to end against the golden values."""
import json
import math
import os
import re
import subprocess

import pytest

from conftest import GOLDEN, ROOT

BIN = os.path.join(ROOT, "openimpala_b200", "bin")


@pytest.fixture(scope="module")
def host_bins(built_lib):
    from openimpala_b200 import build
    build.build_host()
    for exe in ("Diffusion", "tTortuosity", "tReaders", "tEffectiveDiffusivity"):
        assert os.path.exists(os.path.join(BIN, exe))
    return BIN


def run(exe, *args, check=True, env=None):
    r = subprocess.run([os.path.join(BIN, exe), *args], cwd=ROOT, capture_output=True, text=True, timeout=600,
                       env=None if env is None else {**os.environ, **env})
    if check:
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r


def weighted_checksum(vol01):
    """tReaders' WeightedChecksum of a thresholded volume [z, y, x] (values 0/1)."""
    import numpy as np
    v = np.ascontiguousarray(vol01).reshape(-1).astype(np.uint64)
    w = (np.arange(v.size, dtype=np.uint64) % np.uint64(65521)) + np.uint64(1)
    return int((w * v).sum(dtype=np.uint64))


def field(out, key):
    m = re.search(rf"{key}:\s*(.+)", out)
    assert m, f"{key} not in output"
    return m.group(1).split()


@pytest.mark.parametrize("mode,args,dims,count1", [
    ("tiff", ["tifffile=tests/golden/SampleData_2Phase_stack_3d_1bit.tif"], [100, 100, 100], 398309),
    ("tiff", ["tifffile=tests/golden/spheres.tif"], [100, 100, 100], 888024),      # Photometric=0 is NOT inverted
    ("tiff", ["tifffile=tests/golden/SampleData_2Phase_squared.tif"], [64, 64, 64], 104681),
    ("raw", ["rawfile=tests/golden/SampleData_2Phase_stack_3d_uint8.raw", "width=100", "height=100", "depth=100",
             "datatype=UINT8", "threshold=0.5"], [100, 100, 100], 399553),
    ("hdf5", ["hdf5file=tests/golden/SampleData_2Phase_3d.hdf5", "hdf5dataset=image"], [100, 100, 100], 399553),
])
def test_readers_match_reference_fixtures(host_bins, mode, args, dims, count1):
    r = run("tReaders", f"mode={mode}", "gpu_count=0", *args)
    assert [int(v) for v in field(r.stdout, "Dims")] == dims
    assert field(r.stdout, "ThresholdMinMax") == ["0", "1"]
    assert int(field(r.stdout, "DirectCount1")[0]) == count1
    if mode == "tiff" and "1bit" in args[0]:
        assert "BitsPerSample: 1 SampleFormat: 1 SamplesPerPixel: 1" in r.stdout   # tTiffReader.cpp
    if mode in ("raw", "hdf5"):
        # the HDF5 sample's payload is the RAW sample's bytes: same voxels in the same places,
        # through threshold() and through the streamed-upload entry point (11 planes at a time)
        import numpy as np
        raw = np.fromfile(os.path.join(GOLDEN, "SampleData_2Phase_stack_3d_uint8.raw"), dtype=np.uint8).reshape(100, 100, 100)
        assert int(field(r.stdout, "WeightedChecksum")[0]) == weighted_checksum(raw > 0.5)
        r2 = run("tReaders", f"mode={mode}", "gpu_count=0", "u8_chunk=11", *args)
        assert field(r2.stdout, "U8ChunkMismatches") == ["0"]


def test_readers_agree_with_oracle_decoder(host_bins):
    from oracle import oi_numpy as o
    for name in ("SampleData_2Phase_stack_3d_1bit.tif", "spheres.tif", "SampleData_2Phase_squared.tif"):
        ph = o.threshold(o.read_tiff_raw(os.path.join(GOLDEN, name)))
        r = run("tReaders", "mode=tiff", "gpu_count=0", f"tifffile=tests/golden/{name}", "u8_chunk=7")
        assert field(r.stdout, "U8ChunkMismatches") == ["0"]         # streamed-upload entry point, 7 planes at a time
        assert int(field(r.stdout, "DirectCount1")[0]) == int(ph.sum())
        assert int(field(r.stdout, "WeightedChecksum")[0]) == weighted_checksum(ph)


@pytest.mark.parametrize("compression", ["tiff_lzw", "tiff_adobe_deflate", "packbits"])
@pytest.mark.parametrize("kind", ["u8", "u16", "bit", "u8_pred"])
def test_compressed_tiff_stacks(host_bins, tmp_path, compression, kind):
    """Compressed stacks (libtiff decodes them for the reference, src/io/TiffReader.cpp:289-444):
    written with Pillow, read back by the host reader, counted against numpy."""
    import numpy as np
    from PIL import Image
    rng = np.random.default_rng(17)
    nz, ny, nx = 5, 37, 70                                   # not multiples of 8: partial bytes / strips
    base = (rng.random((nz, ny, nx)) < 0.4)
    base[:, 10:20, :] = True                                 # long runs so the coders really compress
    extra = {}
    if kind == "u8":
        vol, thr = (base * 200 + rng.integers(0, 50, base.shape)).astype(np.uint8), 100
        pages = [Image.fromarray(p) for p in vol]
    elif kind == "u8_pred":
        if compression == "packbits":
            pytest.skip("predictor applies to LZW / Deflate only")
        vol, thr = (base * 200 + rng.integers(0, 50, base.shape)).astype(np.uint8), 100
        pages = [Image.fromarray(p) for p in vol]
        extra = {"tiffinfo": {317: 2}}
    elif kind == "u16":
        vol, thr = (base * 30000 + rng.integers(0, 5000, base.shape)).astype(np.uint16), 20000
        pages = [Image.fromarray(p) for p in vol]
    else:
        vol, thr = base.astype(np.uint8), 0.5
        pages = [Image.fromarray(p * 255).convert("1") for p in vol]
    f = tmp_path / f"stack_{kind}_{compression}.tif"
    pages[0].save(f, save_all=True, append_images=pages[1:], compression=compression, **extra)
    r = run("tReaders", "mode=tiff", "gpu_count=0", f"tifffile={f}", f"threshold={thr}", "u8_chunk=3")
    assert [int(v) for v in field(r.stdout, "Dims")] == [nx, ny, nz]
    assert int(field(r.stdout, "DirectCount1")[0]) == int((vol > thr).sum())
    assert int(field(r.stdout, "WeightedChecksum")[0]) == weighted_checksum(vol > thr)
    assert field(r.stdout, "U8ChunkMismatches") == ["0"]


def _write_tiff(path, vol, *, big=False, little=True, tile=None, rows_per_strip=None):
    """Minimal TIFF writer for the reader tests: classic or BigTIFF, either byte order, strips
    or tiles, one IFD per z-plane, uint8 / uint16 / float32 samples."""
    import struct
    import numpy as np
    bo = "<" if little else ">"
    fmt = {np.dtype("uint8"): 1, np.dtype("uint16"): 1, np.dtype("float32"): 3}[vol.dtype]
    bps = vol.dtype.itemsize * 8
    nz, ny, nx = vol.shape
    data = vol.astype(vol.dtype.newbyteorder(bo))
    out = bytearray()
    out += (b"II" if little else b"MM")
    out += struct.pack(bo + "H", 43 if big else 42)
    out += struct.pack(bo + "HHQ", 8, 0, 0) if big else struct.pack(bo + "I", 0)
    first_off_pos = 8 if big else 4
    prev_next_pos = first_off_pos
    for k in range(nz):
        plane = data[k]
        segs = []
        if tile:
            tw, th = tile
            for oy in range(0, ny, th):
                for ox in range(0, nx, tw):
                    t = np.zeros((th, tw), dtype=plane.dtype)
                    blk = plane[oy:oy + th, ox:ox + tw]
                    t[:blk.shape[0], :blk.shape[1]] = blk
                    segs.append(t.tobytes())
        else:
            rps = rows_per_strip or ny
            for oy in range(0, ny, rps):
                segs.append(plane[oy:oy + rps].tobytes())
        offs = []
        for sg in segs:
            if len(out) % 2:
                out += b"\0"
            offs.append(len(out))
            out += sg
        cnts = [len(sg) for sg in segs]
        tags = [(256, nx), (257, ny), (258, bps), (259, 1), (262, 1), (277, 1), (284, 1), (339, fmt)]
        arrays = {}
        if tile:
            tags += [(322, tile[0]), (323, tile[1])]
            arrays[324], arrays[325] = offs, cnts
        else:
            tags += [(278, rows_per_strip or ny)]
            arrays[273], arrays[279] = offs, cnts
        # array payloads (LONG8 in BigTIFF, LONG otherwise)
        arr_pos = {}
        for tag, vals in arrays.items():
            if len(out) % 2:
                out += b"\0"
            arr_pos[tag] = len(out)
            out += struct.pack(bo + ("Q" if big else "I") * len(vals), *vals)
        if len(out) % 2:
            out += b"\0"
        ifd_pos = len(out)
        entries = []
        for tag, v in tags:
            entries.append((tag, 4, 1, v, None))                     # LONG scalar
        for tag, vals in arrays.items():
            entries.append((tag, 16 if big else 4, len(vals), vals, arr_pos[tag]))
        entries.sort(key=lambda e: e[0])
        out += struct.pack(bo + ("Q" if big else "H"), len(entries))
        inline = 8 if big else 4
        for tag, typ, cnt, v, pos in entries:
            out += struct.pack(bo + "HH" + ("Q" if big else "I"), tag, typ, cnt)
            size = (8 if typ == 16 else 4) * cnt
            if pos is not None and size > inline:
                out += struct.pack(bo + ("Q" if big else "I"), pos)
            else:
                vals = v if isinstance(v, list) else [v]
                raw = struct.pack(bo + ("Q" if typ == 16 else "I") * cnt, *vals)
                out += raw + b"\0" * (inline - len(raw))
        next_pos = len(out)
        out += struct.pack(bo + ("Q" if big else "I"), 0)
        struct.pack_into(bo + ("Q" if big else "I"), out, prev_next_pos, ifd_pos)
        prev_next_pos = next_pos
    open(path, "wb").write(bytes(out))


@pytest.mark.parametrize("variant", ["classic_le_strips", "classic_be_strips", "bigtiff_le_tiles", "classic_be_tiles",
                                     "bigtiff_be_strips_f32", "classic_le_multi_strip_u16"])
def test_tiff_container_variants(host_bins, tmp_path, variant):
    """Byte orders, classic / BigTIFF containers, strips / tiles, integer and float samples
    (what libtiff gives the reference for free, src/io/TiffReader.cpp:289-444)."""
    import numpy as np
    rng = np.random.default_rng(23)
    nz, ny, nx = 4, 23, 41
    if "f32" in variant:
        vol, thr = rng.random((nz, ny, nx)).astype(np.float32), 0.6
    elif "u16" in variant:
        vol, thr = rng.integers(0, 60000, (nz, ny, nx)).astype(np.uint16), 30000
    else:
        vol, thr = rng.integers(0, 255, (nz, ny, nx)).astype(np.uint8), 120
    f = tmp_path / f"{variant}.tif"
    _write_tiff(f, vol, big="bigtiff" in variant, little="_le_" in variant,
                tile=(16, 16) if "tiles" in variant else None,
                rows_per_strip=5 if "multi_strip" in variant else None)
    r = run("tReaders", "mode=tiff", "gpu_count=0", f"tifffile={f}", f"threshold={thr}", "u8_chunk=3")
    assert [int(v) for v in field(r.stdout, "Dims")] == [nx, ny, nz]
    assert int(field(r.stdout, "DirectCount1")[0]) == int((vol > thr).sum())
    assert int(field(r.stdout, "WeightedChecksum")[0]) == weighted_checksum(vol > thr)
    assert field(r.stdout, "U8ChunkMismatches") == ["0"]


def test_tiff_file_sequence(host_bins, tmp_path):
    # numbered single-plane files: base pattern + zero-padded index + suffix (TiffReader.H:63-120)
    import numpy as np
    rng = np.random.default_rng(29)
    vol = rng.integers(0, 255, (6, 12, 18)).astype(np.uint8)
    for k in range(6):
        _write_tiff(tmp_path / f"slice_{k + 3:04d}.tif", vol[k:k + 1])
    r = run("tReaders", "mode=tiffseq", "gpu_count=0", f"tifffile={tmp_path / 'slice_'}", "num_files=6",
            "start_index=3", "digits=4", "threshold=99", "u8_chunk=4")
    assert field(r.stdout, "U8ChunkMismatches") == ["0"]
    assert [int(v) for v in field(r.stdout, "Dims")] == [18, 12, 6]
    assert int(field(r.stdout, "DirectCount1")[0]) == int((vol > 99).sum())
    assert int(field(r.stdout, "WeightedChecksum")[0]) == weighted_checksum(vol > 99)


def test_tiff_inconsistent_pages_abort(host_bins, tmp_path):
    """A later directory / sequence file that differs from the first one (size, bit depth, compression)
    must abort instead of being thresholded with the first page's metadata (the reference re-queries
    libtiff per directory, src/io/TiffReader.cpp:320-352)."""
    import struct
    import numpy as np
    rng = np.random.default_rng(31)
    a = rng.integers(0, 255, (1, 12, 18)).astype(np.uint8)
    # sequence: second file has another size; third case: another bit depth
    _write_tiff(tmp_path / "s_0000.tif", a)
    _write_tiff(tmp_path / "s_0001.tif", rng.integers(0, 255, (1, 12, 20)).astype(np.uint8))
    r = run("tReaders", "mode=tiffseq", "gpu_count=0", f"tifffile={tmp_path / 's_'}", "num_files=2", "start_index=0",
            "digits=4", "threshold=99", check=False)
    assert r.returncode != 0 and "inconsistent TIFF directory" in (r.stdout + r.stderr)
    _write_tiff(tmp_path / "t_0000.tif", a)
    _write_tiff(tmp_path / "t_0001.tif", rng.integers(0, 60000, (1, 12, 18)).astype(np.uint16))
    r = run("tReaders", "mode=tiffseq", "gpu_count=0", f"tifffile={tmp_path / 't_'}", "num_files=2", "start_index=0",
            "digits=4", "threshold=99", check=False)
    assert r.returncode != 0 and "inconsistent TIFF directory" in (r.stdout + r.stderr)
    # stack: patch the Compression tag (259) of the second directory to JPEG (7)
    f = tmp_path / "stack.tif"
    _write_tiff(f, rng.integers(0, 255, (3, 12, 18)).astype(np.uint8))
    raw = bytearray(open(f, "rb").read())
    off = struct.unpack_from("<I", raw, 4)[0]
    nent = struct.unpack_from("<H", raw, off)[0]
    off = struct.unpack_from("<I", raw, off + 2 + 12 * nent)[0]          # second IFD
    nent = struct.unpack_from("<H", raw, off)[0]
    for e in range(nent):
        if struct.unpack_from("<H", raw, off + 2 + 12 * e)[0] == 259:
            struct.pack_into("<I", raw, off + 2 + 12 * e + 8, 7)
    open(f, "wb").write(bytes(raw))
    r = run("tReaders", "mode=tiff", "gpu_count=0", f"tifffile={f}", "threshold=99", check=False)
    assert r.returncode != 0 and "inconsistent TIFF directory" in (r.stdout + r.stderr)


@pytest.mark.parametrize("datatype,dtype,thr", [
    ("UINT8", "u1", 100.5), ("INT16_LE", "<i2", -3.5), ("UINT16_LE", "<u2", 30000.0), ("UINT16_BE", ">u2", 30000.0),
    ("FLOAT32_LE", "<f4", 0.25),
])
@pytest.mark.parametrize("threads", ["1", "5"])
def test_raw_reader_types_and_threads(host_bins, tmp_path, datatype, dtype, thr, threads, monkeypatch):
    """RawReader::threshold for the 8/16-bit lookup-table paths and the generic per-sample path,
    single- and multi-threaded (OI_IO_THREADS), against numpy: (double(v) > thr) ? 1 : 0
    (reference src/io/RawReader.cpp:379-491)."""
    import numpy as np
    rng = np.random.default_rng(17)
    shape = (9, 11, 13)                                           # z, y, x
    if dtype.endswith("f4"):
        a = rng.standard_normal(shape).astype(dtype)
    else:
        info = np.iinfo(np.dtype(dtype))
        a = rng.integers(info.min, info.max, size=shape, endpoint=True).astype(dtype)
    path = tmp_path / "vol.raw"
    a.tofile(path)
    monkeypatch.setenv("OI_IO_THREADS", threads)
    r = run("tReaders", "mode=raw", f"rawfile={path}", "width=13", "height=11", "depth=9", f"datatype={datatype}",
            f"threshold={thr}", "gpu_count=0", "u8_chunk=4")
    assert "TEST PASSED" in r.stdout
    assert field(r.stdout, "U8ChunkMismatches") == ["0"]
    assert field(r.stdout, "Dims") == ["13", "11", "9"]
    assert int(field(r.stdout, "DirectCount1")[0]) == int((a.astype(np.float64) > thr).sum())
    assert int(field(r.stdout, "WeightedChecksum")[0]) == weighted_checksum(a.astype(np.float64) > thr)


def _write_hdf5(path, name, vol, chunks=None, shuffle=False, gzip=False, fletcher=False, skip_chunk=None,
                unfiltered_chunk=None, leaf_fanout=5):
    """A version-0-superblock HDF5 file with one old-style root group and one rank-3 dataset
    `name`, written from the HDF5 file format specification (no h5py in this image):
    contiguous, or chunked with a (two-level) version-1 B-tree chunk index and the shuffle /
    deflate / fletcher32 filters.  skip_chunk: index of a chunk that is never allocated (reads
    as the fill value 0); unfiltered_chunk: index of a chunk stored with its filter-mask bits
    set (raw bytes)."""
    import struct
    import zlib
    import numpy as np
    UNDEF = 0xFFFFFFFFFFFFFFFF
    dt = vol.dtype
    es = dt.itemsize
    big = dt.byteorder == ">"
    buf = bytearray()

    def pad8(b):
        return b + b"\0" * (-len(b) % 8)

    def msg(mtype, data):
        data = pad8(data)
        return struct.pack("<HHB3x", mtype, len(data), 0) + data

    def obj_header(msgs):
        body = b"".join(msgs)
        return struct.pack("<BxHII4x", 1, len(msgs), 1, len(body)) + body

    # ---- fixed front matter: superblock (96 bytes), root header, heap, group B-tree, SNOD
    name_b = name.encode() + b"\0"
    heap_data = pad8(b"\0") + pad8(name_b)
    name_off = 8
    heap_data += struct.pack("<QQ", 1, 16)                        # one free block (next = none, 16 bytes)
    ROOT_HDR = 96
    root_hdr_len = 16 + 8 + 16
    HEAP = ROOT_HDR + root_hdr_len
    HEAP_DATA = HEAP + 32
    GTREE = HEAP_DATA + len(heap_data)
    gtree_len = 8 + 16 + 8 + 8 + 8
    SNOD = GTREE + gtree_len
    snod_len = 8 + 40
    DSET = SNOD + snod_len

    # ---- dataset messages
    nz, ny, nx = vol.shape
    dataspace = struct.pack("<BBB5x", 1, 3, 0) + struct.pack("<QQQ", nz, ny, nx)
    if dt.kind == "f":
        bits0 = (1 if big else 0) | 0x20
        dtype_msg = struct.pack("<BBBBI", 0x11, bits0, es * 8 - 1, 0, es)
        dtype_msg += (struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127) if es == 4 else struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023))
    else:
        bits0 = (1 if big else 0) | (0x08 if dt.kind == "i" else 0)
        dtype_msg = struct.pack("<BBBBI", 0x10, bits0, 0, 0, es) + struct.pack("<HH", 0, es * 8)
    filters = []
    if shuffle:
        filters.append((2, [es]))
    if gzip:
        filters.append((1, [4]))
    if fletcher:
        filters.append((3, []))
    msgs_wo_layout = [msg(0x0001, dataspace), msg(0x0003, dtype_msg)]
    if filters:
        fp = struct.pack("<BB6x", 1, len(filters))
        for fid, client in filters:
            fp += struct.pack("<HHHH", fid, 0, 0, len(client)) + b"".join(struct.pack("<I", c) for c in client)
            if len(client) & 1:
                fp += b"\0\0\0\0"
        msgs_wo_layout.append(msg(0x000B, fp))
    layout_len = len(msg(0x0008, b"\0" * (27 if chunks else 18)))
    dset_len = 16 + sum(len(m) for m in msgs_wo_layout) + layout_len
    DATA = DSET + dset_len                                        # first byte after the dataset header

    tail = bytearray()                                            # everything from DATA on
    if not chunks:
        raw = np.ascontiguousarray(vol).tobytes()
        layout = struct.pack("<BBQQ", 3, 1, DATA, len(raw))
        tail += raw
    else:
        cz, cy, cx = chunks
        entries = []                                              # (size, mask, (z, y, x), address)
        idx = 0
        for z0 in range(0, nz, cz):
            for y0 in range(0, ny, cy):
                for x0 in range(0, nx, cx):
                    block = np.zeros((cz, cy, cx), dtype=dt)
                    part = vol[z0:z0 + cz, y0:y0 + cy, x0:x0 + cx]
                    block[:part.shape[0], :part.shape[1], :part.shape[2]] = part
                    data = block.tobytes()
                    mask = 0
                    if idx == unfiltered_chunk:
                        mask = (1 << len(filters)) - 1
                    else:
                        for fid, _ in filters:
                            if fid == 2 and es > 1:
                                a = np.frombuffer(data, dtype=np.uint8).reshape(-1, es)
                                data = np.ascontiguousarray(a.T).tobytes()
                            elif fid == 1:
                                data = zlib.compress(data, 4)
                            elif fid == 3:
                                data = data + b"\0\0\0\0"
                    if idx != skip_chunk:
                        entries.append((len(data), mask, (z0, y0, x0), DATA + len(tail)))
                        tail += data
                        tail += b"\0" * (-len(tail) % 8)
                    idx += 1

        def key(size, mask, off):
            return struct.pack("<IIQQQQ", size, mask, off[0], off[1], off[2], 0)

        def node(level, ents):                                    # ents: (size, mask, off, child address)
            b = b"TREE" + struct.pack("<BBHQQ", 1, level, len(ents), UNDEF, UNDEF)
            for size, mask, off, child in ents:
                b += key(size, mask, off) + struct.pack("<Q", child)
            return b + key(0, 0, (nz, ny, nx))                    # closing key
        leaves = [entries[i:i + leaf_fanout] for i in range(0, len(entries), leaf_fanout)]
        uppers = []
        for leaf in leaves:
            addr = DATA + len(tail)
            tail += node(0, leaf)
            uppers.append((leaf[0][0], leaf[0][1], leaf[0][2], addr))
        if len(uppers) == 1:
            btree = uppers[0][3]
        elif uppers:
            btree = DATA + len(tail)
            tail += node(1, uppers)
        else:
            btree = UNDEF
        layout = struct.pack("<BBBQIIII", 3, 2, 4, btree, cz, cy, cx, es)
    dset = obj_header(msgs_wo_layout + [msg(0x0008, layout)])
    assert len(dset) == dset_len

    eof = DATA + len(tail)
    sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII16x", 0, ROOT_HDR, 0, 0)             # root group symbol table entry
    assert len(sb) == 96
    root = obj_header([msg(0x0011, struct.pack("<QQ", GTREE, HEAP))])
    assert len(root) == root_hdr_len
    heap = b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), len(heap_data) - 16, HEAP_DATA)
    gtree = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF) + struct.pack("<QQQ", 0, SNOD, name_off)
    assert len(gtree) == gtree_len
    snod = b"SNOD" + struct.pack("<BxH", 1, 1) + struct.pack("<QQII16x", name_off, DSET, 0, 0)
    assert len(snod) == snod_len
    buf += sb + root + heap + heap_data + gtree + snod + dset + tail
    assert len(buf) == eof
    with open(path, "wb") as fh:
        fh.write(buf)


@pytest.mark.parametrize("case", ["contiguous_u8", "chunked_u8", "chunked_u8_gzip", "chunked_u16_shuffle_gzip",
                                  "chunked_i16be_gzip_fletcher", "chunked_f32_shuffle_gzip_partial", "chunked_sparse"])
@pytest.mark.parametrize("threads", ["1", "6"])
def test_hdf5_chunked_and_filtered(host_bins, tmp_path, case, threads, monkeypatch):
    """Chunked HDF5 datasets through the version-1 B-tree chunk index with the deflate, shuffle
    and fletcher32 filters (what libhdf5 resolves for the reference, src/io/HDF5Reader.cpp:255-402),
    edge chunks that overhang the dataset, a chunk stored unfiltered (filter mask), a chunk that
    was never allocated (fill value) and a two-level B-tree; files written from the format
    specification by _write_hdf5 (its contiguous output goes through the code path the
    reference's real sample file exercises)."""
    import numpy as np
    rng = np.random.default_rng(41)
    shape = (11, 13, 17)                                          # z, y, x: no extent is a multiple of the chunk's
    kw = {}
    if case.endswith("_u8") or "u8_gzip" in case or case == "chunked_sparse":
        vol, thr = rng.integers(0, 255, shape).astype("u1"), 120.5
    elif "u16" in case:
        vol, thr = rng.integers(0, 60000, shape).astype("<u2"), 30000.0
    elif "i16be" in case:
        vol, thr = rng.integers(-20000, 20000, shape).astype(">i2"), -100.5
    else:
        vol, thr = rng.standard_normal(shape).astype("<f4"), 0.1
    if case != "contiguous_u8":
        kw["chunks"] = (4, 5, 6)
    kw["gzip"] = "gzip" in case
    kw["shuffle"] = "shuffle" in case
    kw["fletcher"] = "fletcher" in case
    expect = vol.astype(np.float64)
    if "partial" in case:
        kw["unfiltered_chunk"] = 7
    if case == "chunked_sparse":
        kw["gzip"] = True
        kw["skip_chunk"] = 10                                     # 3 x 3 x 3 chunks: index 10 = offsets (4, 0, 6)
        expect = expect.copy()
        expect[4:8, 0:5, 6:12] = 0.0
    f = tmp_path / f"{case}.h5"
    _write_hdf5(f, "image", vol, **kw)
    monkeypatch.setenv("OI_IO_THREADS", threads)
    r = run("tReaders", "mode=hdf5", f"hdf5file={f}", "hdf5dataset=image", f"threshold={thr}", "gpu_count=0", "u8_chunk=3")
    assert "TEST PASSED" in r.stdout, r.stdout + r.stderr
    assert [int(v) for v in field(r.stdout, "Dims")] == [17, 13, 11]
    assert int(field(r.stdout, "DirectCount1")[0]) == int((expect > thr).sum())
    assert int(field(r.stdout, "WeightedChecksum")[0]) == weighted_checksum(expect > thr)
    assert field(r.stdout, "U8ChunkMismatches") == ["0"]


def test_dat_reader(host_bins, tmp_path):
    # src/io/DatReader.cpp:60-248: int32 LE (W, H, D) header + uint16 LE voxels, x fastest
    import numpy as np
    rng = np.random.default_rng(9)
    vol = rng.integers(0, 4000, size=(5, 7, 9), dtype=np.uint16)          # [z, y, x]
    f = tmp_path / "vol.dat"
    with open(f, "wb") as fh:
        fh.write(np.array([9, 7, 5], dtype="<i4").tobytes())
        fh.write(vol.astype("<u2").tobytes())
        fh.write(b"extra")                                                # trailing bytes are ignored with a warning
    r = run("tReaders", "mode=dat", "gpu_count=0", f"datfile={f}", "threshold=1999")
    assert [int(v) for v in field(r.stdout, "Dims")] == [9, 7, 5]
    assert int(field(r.stdout, "DirectCount1")[0]) == int((vol > 1999).sum())
    assert int(field(r.stdout, "WeightedChecksum")[0]) == weighted_checksum(vol > 1999)
    assert [int(v) for v in field(r.stdout, "RawCorner")] == [int(vol[0, 0, 0]), int(vol[-1, -1, -1])]
    # truncated payload and bad header are rejected
    (tmp_path / "short.dat").write_bytes(np.array([9, 7, 5], dtype="<i4").tobytes() + b"\0" * 10)
    assert run("tReaders", "mode=dat", "gpu_count=0", f"datfile={tmp_path / 'short.dat'}", check=False).returncode != 0
    (tmp_path / "bad.dat").write_bytes(np.array([0, 7, 5], dtype="<i4").tobytes())
    assert run("tReaders", "mode=dat", "gpu_count=0", f"datfile={tmp_path / 'bad.dat'}", check=False).returncode != 0


def read_amrex_plotfile(path):
    """Independent reader of a single-level, single-grid AMReX plotfile (layout of AMReX 25.03
    WriteGenericPlotfileHeader + VisMF header version 1): returns (header dict, {name: array[z,y,x]})."""
    import numpy as np
    lines = open(os.path.join(path, "Header")).read().split("\n")
    it = iter(lines)
    h = {"version": next(it)}
    ncomp = int(next(it))
    h["names"] = [next(it) for _ in range(ncomp)]
    h["dim"] = int(next(it))
    h["time"] = float(next(it))
    h["finest_level"] = int(next(it))
    h["prob_lo"] = [float(v) for v in next(it).split()]
    h["prob_hi"] = [float(v) for v in next(it).split()]
    h["ref_ratio"] = next(it).split()
    m = re.fullmatch(r"\(\((\d+),(\d+),(\d+)\) \((\d+),(\d+),(\d+)\) \(0,0,0\)\) ?", next(it))
    assert m, "domain box line"
    h["domain"] = [int(v) for v in m.groups()]
    h["level_steps"] = [int(v) for v in next(it).split()]
    h["dx"] = [float(v) for v in next(it).split()]
    h["coord_sys"] = int(next(it))
    h["bwidth"] = int(next(it))
    lev, ngrids, t = next(it).split()
    h["grids"] = int(ngrids)
    assert int(lev) == 0 and float(t) == h["time"]
    h["step"] = int(next(it))
    h["grid_extent"] = [[float(v) for v in next(it).split()] for _ in range(3)]
    h["mf_path"] = next(it)
    # VisMF header
    ch = open(os.path.join(path, h["mf_path"] + "_H")).read().split("\n")
    it = iter(ch)
    assert int(next(it)) == 1 and int(next(it)) == 1          # version 1, one file per process
    assert int(next(it)) == ncomp
    assert int(next(it)) == 0                                 # no ghost cells
    assert next(it) == "(1 0"
    box_line = next(it)
    assert next(it) == ")"
    assert int(next(it)) == 1
    fod = next(it).split()
    assert fod[0] == "FabOnDisk:" and int(fod[2]) == 0
    assert next(it) == ""
    assert next(it) == f"1,{ncomp}"
    h["min"] = [float(v) for v in next(it).rstrip(",").split(",")]
    assert next(it) == ""
    assert next(it) == f"1,{ncomp}"
    h["max"] = [float(v) for v in next(it).rstrip(",").split(",")]
    raw = open(os.path.join(path, os.path.dirname(h["mf_path"]), fod[1]), "rb").read()
    nl = raw.index(b"\n")
    fab = raw[:nl].decode()
    assert fab.startswith("FAB ((8, (64 11 52 0 1 12 0 1023)),(8, (8 7 6 5 4 3 2 1)))")   # native little-endian IEEE double
    assert fab.endswith(box_line + f" {ncomp}")
    lo_hi = h["domain"]
    nx, ny, nz = (lo_hi[3 + d] - lo_hi[d] + 1 for d in range(3))
    data = np.frombuffer(raw[nl + 1:], dtype="<f8")
    assert data.size == ncomp * nx * ny * nz
    data = data.reshape(ncomp, nz, ny, nx)
    return h, {n: data[c] for c, n in enumerate(h["names"])}


def test_plotfile_writer_round_trip(host_bins, tmp_path):
    """amrex::WriteSingleLevelPlotfile of the stand-in (the writer behind write_plotfile = 1,
    reference TortuosityHypre.cpp:710-745): Header / Cell_H / FAB parse back with an independent
    reader and carry the fields bit for bit."""
    import numpy as np
    from oracle import oi_numpy as o
    plt = str(tmp_path / "plt_check")
    r = run("tReaders", "mode=tiff", "tifffile=tests/golden/SampleData_2Phase_squared.tif", "gpu_count=0",
            f"plotfile={plt}")
    assert "TEST PASSED" in r.stdout
    h, f = read_amrex_plotfile(plt)
    assert h["version"] == "HyperCLaw-V1.1" and h["names"] == ["phase_id", "ramp"] and h["dim"] == 3
    assert h["domain"] == [0, 0, 0, 63, 63, 63] and h["finest_level"] == 0 and h["grids"] == 1
    assert h["prob_lo"] == [0.0, 0.0, 0.0] and h["prob_hi"] == [64.0, 64.0, 64.0] and h["dx"] == [1.0, 1.0, 1.0]
    assert h["grid_extent"] == [[0.0, 64.0]] * 3 and h["mf_path"] == "Level_0/Cell"
    ph = o.threshold(o.read_tiff_raw(os.path.join(GOLDEN, "SampleData_2Phase_squared.tif")), 0.5)
    assert np.array_equal(f["phase_id"], ph.astype(np.float64))
    k, j, i = np.meshgrid(np.arange(64), np.arange(64), np.arange(64), indexing="ij")
    assert np.array_equal(f["ramp"], i + 0.5 * j - 0.25 * k)
    assert h["min"] == [0.0, f["ramp"].min()] and h["max"] == [1.0, f["ramp"].max()]


def test_missing_required_key_aborts(host_bins):
    r = run("Diffusion", "calculation_method=flow_through", check=False)
    assert r.returncode != 0 and "filename" in r.stderr


def test_unknown_method_aborts(host_bins):
    r = run("Diffusion", "filename=SampleData_2Phase_squared.tif", "data_path=tests/golden/",
            "results_path=gpurun_out/r0/", "calculation_method=nonsense", check=False)
    assert r.returncode != 0 and "Invalid calculation_method" in r.stderr


def test_no_gpu_aborts_loudly(host_bins):
    from openimpala_b200 import capi
    if capi.device_count() > 0:
        pytest.skip("a GPU is present")
    r = run("tReaders", "mode=tiff", "tifffile=tests/golden/SampleData_2Phase_squared.tif", check=False)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


# ------------------------------------------------------------------ GPU: the apps end to end
@pytest.mark.gpu
def test_tvolumefraction_counts_on_gpu(host_bins):
    r = run("tReaders", "mode=tiff", "tifffile=tests/golden/SampleData_2Phase_stack_3d_1bit.tif")
    assert "VolumeFractionCount1: 398309 of 1000000" in r.stdout
    assert "VolumeFractionCount0: 601691 of 1000000" in r.stdout
    assert "TEST PASSED" in r.stdout


@pytest.mark.gpu
def test_ttortuosity_driver(host_bins):
    r = run("tTortuosity", "tests/inputs/tTortuosity.inputs")
    assert "Matrix/vector property checks passed" in r.stdout
    assert "Conservation Check Status: PASS" in r.stdout
    assert "TEST PASSED" in r.stdout
    tau = float(field(r.stdout, "Final Calculated Tortuosity")[0])
    assert abs(tau - 1.6934074851) <= 1e-6 * 1.6934074851


@pytest.mark.gpu
def test_diffusion_app_results_txt(host_bins):
    gold = json.load(open(os.path.join(GOLDEN, "sample_golden.json")))
    run("Diffusion", "tests/inputs/diffusion_flow_through.inputs")
    txt = open(os.path.join(ROOT, "gpurun_out", "results_diffusion", "results.txt")).read()
    lines = [l for l in txt.splitlines() if not l.startswith("#")]
    vals = dict(l.split(": ") for l in lines)
    assert list(vals) == ["VolumeFraction", "Tortuosity_X", "Tortuosity_Y", "Tortuosity_Z"]   # map-sorted
    assert vals["VolumeFraction"] == "0.398309000"                                            # %.9f
    for d, key in enumerate(["Tortuosity_X", "Tortuosity_Y", "Tortuosity_Z"]):
        ref = next(c for c in gold["cases"] if c["phase"] == 1 and c["direction"] == d)["tau"]
        assert abs(float(vals[key]) - ref) <= 1e-6 * ref
        assert re.fullmatch(r"\d+\.\d{9}", vals[key])


@pytest.mark.gpu
@pytest.mark.parametrize("ranks", [2, 4])
def test_diffusion_app_on_several_ranks(host_bins, ranks, tmp_path):
    """`Diffusion inputs b200.ranks=N`: the app starts N ranks (one per GPU), every rank reads its own z-slab of the
    TIFF, the C++ classes take the slab from BoxArray / DistributionMapping and the rank's communicator
    (reference: SPMD over MPI ranks, TortuosityHypre.H:68-80, Diffusion.cpp:266-268).  results.txt must equal the
    one-rank file: VolumeFraction identical, tau to 1e-8."""
    from openimpala_b200 import capi
    if capi.device_count() < ranks:
        pytest.skip(f"needs >= {ranks} GPUs")
    out1, outn = tmp_path / "one", tmp_path / "many"
    run("Diffusion", "tests/inputs/diffusion_flow_through.inputs", f"results_path={out1}")
    run("Diffusion", "tests/inputs/diffusion_flow_through.inputs", f"results_path={outn}", f"b200.ranks={ranks}")

    def parse(path):
        lines = [l for l in open(os.path.join(path, "results.txt")).read().splitlines() if not l.startswith("#")]
        return dict(l.split(": ") for l in lines)
    a, b = parse(out1), parse(outn)
    assert list(a) == list(b) == ["VolumeFraction", "Tortuosity_X", "Tortuosity_Y", "Tortuosity_Z"]
    assert a["VolumeFraction"] == b["VolumeFraction"] == "0.398309000"
    for k in ("Tortuosity_X", "Tortuosity_Y", "Tortuosity_Z"):
        assert abs(float(a[k]) - float(b[k])) <= 1e-8 * float(a[k]), (k, a[k], b[k])


@pytest.mark.gpu
def test_diffusion_homogenization_on_two_ranks(host_bins, tmp_path):
    """The app's default method on z-slabs: `EffectiveDiffusivityHypre` takes this rank's slab from the BoxArray
    and the rank's communicator (reference: SPMD over MPI ranks, EffectiveDiffusivityHypre.H:55-63); the periodic wrap
    in z between the last and the first slab and the reduction of the gradient sums are the library's.  The D_eff
    tensor of two ranks (ragged slabs of 64 + 36 planes) must equal the one-rank tensor."""
    from openimpala_b200 import capi
    if capi.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    common = ("filename=SampleData_2Phase_stack_3d_1bit.tif", "data_path=tests/golden/", "phase_id=1", "verbose=0")
    out1, out2 = tmp_path / "one", tmp_path / "two"
    run("Diffusion", *common, f"results_path={out1}")
    run("Diffusion", *common, f"results_path={out2}", "b200.ranks=2")

    def parse(path):
        txt = open(os.path.join(path, "results.txt")).read()
        return {k: float(v) for k, v in re.findall(r"^(Deff_[xyz][xyz]): (\S+)$", txt, flags=re.M)}
    a, b = parse(out1), parse(out2)
    assert len(a) == 9 and list(a) == list(b)
    gold = json.load(open(os.path.join(GOLDEN, "effdiff_golden.json")))["phase1"]["deff"]
    assert abs(a["Deff_xx"] - gold[0][0]) <= 1e-6
    for k in a:
        assert abs(a[k] - b[k]) <= 1e-8, (k, a[k], b[k])


@pytest.mark.gpu
def test_diffusion_write_plotfile(host_bins):
    """write_plotfile = 1 (Diffusion.cpp:211, 696 -> TortuosityHypre.cpp:710-745): the solution,
    the phase ids and the percolation mask as <results_path>/tortuosity_solution_<dir>."""
    import numpy as np
    gold = json.load(open(os.path.join(GOLDEN, "sample_golden.json")))
    run("Diffusion", "tests/inputs/diffusion_flow_through.inputs", "direction=Z", "write_plotfile=1",
        "results_path=gpurun_out/results_plot/", "verbose=0")
    h, f = read_amrex_plotfile(os.path.join(ROOT, "gpurun_out", "results_plot", "tortuosity_solution_2"))
    assert h["names"] == ["solution_potential", "phase_id", "active_mask"]
    assert h["domain"] == [0, 0, 0, 99, 99, 99] and h["prob_hi"] == [100.0, 100.0, 100.0]
    case = next(c for c in gold["cases"] if c["phase"] == 1 and c["direction"] == 2)
    mask = f["active_mask"]
    assert int(mask.sum()) == case["n_active"] and set(np.unique(mask)) <= {0.0, 1.0}
    assert int((f["phase_id"] == 1).sum()) == gold["phase_count"]["1"]
    x = f["solution_potential"]
    act = mask.astype(bool)
    assert not x[~act].any()                                          # zero off the percolating cells
    assert np.all(x[0][act[0]] == -1.0) and np.all(x[-1][act[-1]] == 1.0)     # Dirichlet planes (app defaults vlo/vhi = -1/+1)
    assert x[act].min() >= -1.0 - 1e-6 and x[act].max() <= 1.0 + 1e-6


@pytest.mark.gpu
def test_diffusion_cli_overrides_and_blocked_phase(host_bins):
    # command-line key=value after the inputs file wins (ParmParse), phase 0, one direction
    run("Diffusion", "tests/inputs/diffusion_flow_through.inputs", "phase_id=0", "direction=Z",
        "results_path=gpurun_out/results_p0/", "verbose=0")
    txt = open(os.path.join(ROOT, "gpurun_out", "results_p0", "results.txt")).read()
    assert "# Analysis Phase ID: 0" in txt and "VolumeFraction: 0.601691000" in txt
    tau = float(re.search(r"Tortuosity_Z: (\S+)", txt).group(1))
    assert abs(tau - 1.6930525106) <= 1e-6 * 1.6930525106
    assert "Tortuosity_X" not in txt


@pytest.mark.gpu
def test_diffusion_default_method_is_homogenization(host_bins):
    # the app's default calculation_method (Diffusion.cpp:188, 509-589): three corrector solves + D_eff tensor
    gold = json.load(open(os.path.join(GOLDEN, "effdiff_golden.json")))
    r = run("Diffusion", "filename=SampleData_2Phase_stack_3d_1bit.tif", "data_path=tests/golden/",
            "results_path=gpurun_out/results_homog/", "phase_id=1", "verbose=1", "b200.check_host_tensor=1")
    assert "Full Domain Effective Diffusivity Tensor D_eff / D_material:" in r.stdout
    rows = re.findall(r"^  \[(.+)\]$", r.stdout, flags=re.M)
    D = [[float(v) for v in row.split(",")] for row in rows[-3:]]
    ref = gold["phase1"]["deff"]
    for a in range(3):
        for b in range(3):
            assert abs(D[a][b] - ref[a][b]) <= 1e-6
    assert "Host tensor check" in r.stdout
    txt = open(os.path.join(ROOT, "gpurun_out", "results_homog", "results.txt")).read()
    assert abs(float(re.search(r"Deff_xx: (\S+)", txt).group(1)) - ref[0][0]) <= 1e-6


def _check_rev_study(results_path, env=None):
    # rev.do_study (Diffusion.cpp:317-504): random sub-volumes as periodic boxes -> one CSV row each;
    # every row must equal the oracle's tensor of that very sub-volume
    import numpy as np
    from oracle import oi_effdiff as oe
    from oracle import oi_numpy as o
    run("Diffusion", "filename=SampleData_2Phase_squared.tif", "data_path=tests/golden/",
        f"results_path={results_path}/", "rev.do_study=1", "rev.num_samples=2", "rev.sizes=16 24 200",
        "calculation_method=skip_if_rev", "rev.verbose=0", "verbose=0", env=env)
    lines = open(os.path.join(ROOT, results_path, "rev_study_Deff.csv")).read().splitlines()
    assert lines[0] == ("SampleNo,SeedX,SeedY,SeedZ,REV_Size_Target,ActualSizeX,ActualSizeY,ActualSizeZ,"
                        "D_xx,D_yy,D_zz,D_xy,D_xz,D_yz")
    rows = [l.split(",") for l in lines[1:]]
    assert len(rows) == 6 and [int(r[4]) for r in rows] == [16, 24, 200] * 2
    ph = o.threshold(o.read_tiff_raw(os.path.join(GOLDEN, "SampleData_2Phase_squared.tif")), 0.5)
    for r in rows:
        sx, sy, sz, target = int(r[1]), int(r[2]), int(r[3]), int(r[4])
        ax, ay, az = int(r[5]), int(r[6]), int(r[7])
        assert (ax, ay, az) == ((target,) * 3 if target <= 64 else (64, 64, 64))    # clipped to the domain
        sub = ph[sz:sz + az, sy:sy + ay, sx:sx + ax]
        D = oe.deff_tensor(sub, 1)
        got = [float(v) for v in r[8:14]]
        ref = [D[0][0], D[1][1], D[2][2], D[0][1], D[0][2], D[1][2]]
        assert max(abs(a - b) for a, b in zip(got, ref)) <= 1e-6


@pytest.mark.gpu
def test_teffectivediffusivity_driver(host_bins):
    # reference src/props/tEffectiveDiffusivity.cpp: solves converge, tensor symmetric to 1e-7, diagonals >= 0;
    # plus the device gradient sums against the host's central differences of chi
    import numpy as np
    from oracle import oi_c, oi_numpy as o
    r = run("tEffectiveDiffusivity", "tests/inputs/tEffectiveDiffusivity.inputs")
    assert "TEST RESULT: PASS" in r.stdout and "D_eff tensor symmetry check: PASS" in r.stdout
    ph = o.threshold(o.read_tiff_raw(os.path.join(GOLDEN, "SampleData_2Phase_squared.tif")), 0.5)
    ref = np.asarray(oi_c.effdiff_deff_tensor(ph, 1, eps=1e-10)[0]).reshape(3, 3)
    rows = re.findall(r"\[(\S+), (\S+), (\S+)\]", r.stdout)
    got = np.array([[float(v) for v in row] for row in rows[-3:]])
    assert np.abs(got - ref).max() <= 1e-6


@pytest.mark.gpu
def test_diffusion_rev_study_csv(host_bins):
    _check_rev_study("gpurun_out/results_rev")


def test_host_layer_on_the_mock_rev_study(mock_env, tmp_path):
    # the same driver logic (mt19937 seeding, clipping, CSV) with the device replaced by the mock
    _check_rev_study(str(tmp_path / "rev"), env=mock_env)
    # b200.rev_workers: the sub-volumes on worker threads (replicas) -- same rows in the same order
    outs = []
    for w in (1, 3):
        res = tmp_path / f"w{w}"
        r = run("Diffusion", "filename=SampleData_2Phase_squared.tif", "data_path=tests/golden/", f"results_path={res}/",
                "rev.do_study=1", "rev.num_samples=3", "rev.sizes=12 20 4", "calculation_method=skip_if_rev",
                "rev.verbose=0", "verbose=1", f"b200.rev_workers={w}", env=mock_env)
        assert ("REV workers: 3 host threads" in r.stdout) == (w == 3)
        outs.append(open(res / "rev_study_Deff.csv").read())
    assert outs[0] == outs[1] and len(outs[0].splitlines()) == 1 + 6          # size 4 is skipped (< 8 cells)
    # rev.write_plotfiles: one directory per (sample, size, direction), reference Diffusion.cpp:434-441
    res = tmp_path / "plots"
    run("Diffusion", "filename=SampleData_2Phase_squared.tif", "data_path=tests/golden/", f"results_path={res}/",
        "rev.do_study=1", "rev.num_samples=1", "rev.sizes=12 20", "calculation_method=skip_if_rev", "rev.verbose=0",
        "verbose=0", "rev.write_plotfiles=1", "b200.rev_workers=2", env=mock_env)
    dirs = sorted(d for d in os.listdir(res) if d.startswith("REV_"))
    assert dirs == [f"REV_Sample1_Size{n}_Dir{c}" for n in (12, 20) for c in range(3)]
    h, f = read_amrex_plotfile(str(res / "REV_Sample1_Size20_Dir2" / "effdiff_chi_dir2"))
    assert h["domain"] == [0, 0, 0, 19, 19, 19] and h["names"] == ["chi_k", "active_mask_from_solver"]


@pytest.mark.gpu
def test_diffusion_streamed_tiff_upload(host_bins):
    # b200.stream_upload: planes decoded straight into the pinned staging buffers (no int32 copy)
    gold = json.load(open(os.path.join(GOLDEN, "sample_golden.json")))
    r = run("Diffusion", "tests/inputs/diffusion_flow_through.inputs", "direction=X", "b200.stream_upload=7",
            "results_path=gpurun_out/results_stream/")
    assert "streamed from the TIFF in chunks of 7 planes" in r.stdout
    txt = open(os.path.join(ROOT, "gpurun_out", "results_stream", "results.txt")).read()
    tau = float(re.search(r"Tortuosity_X: (\S+)", txt).group(1))
    ref = next(c for c in gold["cases"] if c["phase"] == 1 and c["direction"] == 0)["tau"]
    assert abs(tau - ref) <= 1e-6 * ref


@pytest.mark.gpu
def test_diffusion_streamed_hdf5_upload(host_bins):
    # the same through HDF5Reader::thresholdPlanesU8 (the reference's HDF5 sample: 399 553 voxels of phase 1)
    r = run("Diffusion", "tests/inputs/diffusion_flow_through.inputs", "filename=SampleData_2Phase_3d.hdf5",
            "direction=Z", "b200.stream_upload=9", "results_path=gpurun_out/results_stream_h5/")
    assert "streamed from the HDF5 dataset in chunks of 9 planes" in r.stdout
    r2 = run("Diffusion", "tests/inputs/diffusion_flow_through.inputs", "filename=SampleData_2Phase_3d.hdf5",
             "direction=Z", "results_path=gpurun_out/results_plain_h5/")
    taus = []
    for d in ("results_stream_h5", "results_plain_h5"):
        txt = open(os.path.join(ROOT, "gpurun_out", d, "results.txt")).read()
        assert "VolumeFraction: 0.399553000" in txt
        taus.append(float(re.search(r"Tortuosity_Z: (\S+)", txt).group(1)))
    assert math.isfinite(taus[0]) and taus[0] > 1.0 and abs(taus[0] - taus[1]) <= 1e-9 * taus[1]


# ------------------------------------------------------------------ host layer on a CPU mock of the C-ABI
@pytest.fixture(scope="module")
def mock_env(host_bins, tmp_path_factory):
    """tests/cpu_emul/mock_capi.c: the C-ABI answered on the CPU by the oracle, LD_PRELOADed in
    front of the real library.  Test infrastructure only -- it lets the readers -> apps -> host
    classes -> results / plotfiles chain run in the GPU-less build container; the product library
    itself has no CPU fallback (test_no_gpu_aborts_loudly)."""
    from oracle import oi_c
    oi_c.load()
    out = tmp_path_factory.mktemp("mock_capi") / "libmock_capi.so"
    odir = os.path.join(ROOT, "oracle")
    subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-o", str(out),
                    os.path.join(ROOT, "tests", "cpu_emul", "mock_capi.c"), "-L", odir, "-l:liboi_oracle.so",
                    "-Wl,-rpath," + odir, "-lm"], check=True)
    return {"LD_PRELOAD": str(out)}


def test_host_layer_on_the_mock_flow_through(mock_env, tmp_path):
    """Diffusion (flow_through, write_plotfile = 1) on the 64^3 sample through readers, VolumeFraction,
    TortuosityHypre::value() and its NaN/flux-gate tail, results.txt and the plotfile -- the device
    replaced by the mock, so tau must be the oracle's."""
    import numpy as np
    from oracle import oi_c, oi_numpy as o
    res = tmp_path / "res"
    r = run("Diffusion", "tests/inputs/diffusion_flow_through.inputs", "filename=SampleData_2Phase_squared.tif",
            "direction=X Z", "write_plotfile=1", f"results_path={res}/", env=mock_env)
    ph = o.threshold(o.read_tiff_raw(os.path.join(GOLDEN, "SampleData_2Phase_squared.tif")), 0.5)
    vals = dict(l.split(": ") for l in open(res / "results.txt").read().splitlines() if not l.startswith("#"))
    assert list(vals) == ["VolumeFraction", "Tortuosity_X", "Tortuosity_Z"]
    assert vals["VolumeFraction"] == f"{ph.mean():.9f}"
    for d, key in ((0, "Tortuosity_X"), (2, "Tortuosity_Z")):
        ref = oi_c.tortuosity(ph, 1, d, -1.0, 1.0, eps=1e-9)
        assert abs(float(vals[key]) - ref["tau"]) <= 1e-6 * ref["tau"]
        h, f = read_amrex_plotfile(str(res / f"tortuosity_solution_{d}"))
        assert h["names"] == ["solution_potential", "phase_id", "active_mask"] and h["domain"] == [0, 0, 0, 63, 63, 63]
        mask = f["active_mask"].astype(bool)
        assert int(mask.sum()) == ref["n_active"]
        assert np.array_equal(f["phase_id"], ph.astype(np.float64))
        x = f["solution_potential"]
        lo = (slice(None), slice(None), 0) if d == 0 else (0, slice(None), slice(None))
        hi = (slice(None), slice(None), -1) if d == 0 else (-1, slice(None), slice(None))
        assert np.all(x[lo][mask[lo]] == -1.0) and np.all(x[hi][mask[hi]] == 1.0) and not x[~mask].any()
    assert "Conservation Check Status: PASS" in r.stdout


def test_host_layer_on_the_mock_direction_workers(mock_env, tmp_path):
    """b200.dir_workers: the three directions as independent jobs on worker threads (one solver
    object, handle and stream each) -- same results.txt, all three plotfiles."""
    outs = []
    for w in (1, 3):
        res = tmp_path / f"w{w}"
        r = run("Diffusion", "tests/inputs/diffusion_flow_through.inputs", "filename=SampleData_2Phase_squared.tif",
                f"results_path={res}/", f"b200.dir_workers={w}", "write_plotfile=1", env=mock_env)
        assert ("Direction workers: 3 host threads" in r.stdout) == (w == 3)
        outs.append(open(res / "results.txt").read())
        assert sorted(os.listdir(res)) == ["results.txt", "tortuosity_solution_0", "tortuosity_solution_1", "tortuosity_solution_2"]
    assert outs[0] == outs[1] and outs[0].count("Tortuosity_") == 3


@pytest.mark.parametrize("kind", ["tiff", "hdf5", "raw"])
def test_host_layer_on_the_mock_streamed_upload(mock_env, tmp_path, kind):
    """b200.stream_upload: chunks decoded straight into the staging buffers give the same tau as
    the iMultiFab path, for TIFF, HDF5 and RAW input."""
    name = {"tiff": "SampleData_2Phase_squared.tif", "hdf5": "SampleData_2Phase_3d.hdf5",
            "raw": "SampleData_2Phase_stack_3d_uint8.raw"}[kind]
    what = {"tiff": "TIFF", "hdf5": "HDF5 dataset", "raw": "RAW volume"}[kind]
    raw_keys = ["width=100", "height=100", "depth=100", "datatype=UINT8"] if kind == "raw" else []
    taus = []
    for extra in (["b200.stream_upload=7"], []):
        res = tmp_path / ("s" if extra else "p")
        r = run("Diffusion", "tests/inputs/diffusion_flow_through.inputs", f"filename={name}", "direction=Y",
                f"results_path={res}/", *raw_keys, *extra, env=mock_env)
        if extra:
            assert f"streamed from the {what} in chunks of 7 planes" in r.stdout
        taus.append(float(re.search(r"Tortuosity_Y: (\S+)", open(res / "results.txt").read()).group(1)))
    assert math.isfinite(taus[0]) and taus[0] == taus[1]


def test_host_layer_on_the_mock_homogenization_and_driver(mock_env, tmp_path):
    """The app's default method (three corrector solves, D_eff tensor, chi plotfiles) and the
    tTortuosity driver, on the mock."""
    import numpy as np
    from oracle import oi_c, oi_numpy as o
    res = tmp_path / "homog"
    run("Diffusion", "tests/inputs/diffusion_flow_through.inputs", "filename=SampleData_2Phase_squared.tif",
        "calculation_method=homogenization", "write_plotfile=1", f"results_path={res}/", env=mock_env)
    txt = open(res / "results.txt").read()
    ph = o.threshold(o.read_tiff_raw(os.path.join(GOLDEN, "SampleData_2Phase_squared.tif")), 0.5)
    d = {k: float(v) for k, v in re.findall(r"(Deff_[xyz][xyz]): (\S+)", txt)}
    assert len(d) == 9
    ref = oi_c.effdiff_deff_tensor(ph, 1, eps=1e-10)
    ref = np.asarray(ref[0] if isinstance(ref, tuple) else ref).reshape(3, 3)
    for ia, a in enumerate("xyz"):
        assert d[f"Deff_{a}{a}"] > 0.0
        for ib, b in enumerate("xyz"):
            assert abs(d[f"Deff_{a}{b}"] - d[f"Deff_{b}{a}"]) <= 1e-7          # tEffectiveDiffusivity.cpp:424-432
            assert abs(d[f"Deff_{a}{b}"] - ref[ia, ib]) <= 1e-6
    h, f = read_amrex_plotfile(str(res / "FullDomain_chi_X" / "effdiff_chi_dir0"))      # reference Diffusion.cpp:531-543
    assert h["names"] == ["chi_k", "active_mask_from_solver"]
    assert np.array_equal(f["active_mask_from_solver"], (ph == 1).astype(np.float64))
    r = run("tTortuosity", "tests/inputs/tTortuosity.inputs", env=mock_env)
    assert "TEST PASSED" in r.stdout or r.returncode == 0
    # the reference's tEffectiveDiffusivity checks: three converged solves, symmetric tensor, diagonals >= 0
    r = run("tEffectiveDiffusivity", "tests/inputs/tEffectiveDiffusivity.inputs", f"resultsdir={tmp_path / 'teff'}",
            "write_plotfile=1", env=mock_env)
    assert "TEST RESULT: PASS" in r.stdout and "D_eff tensor symmetry check: PASS" in r.stdout
    rows = re.findall(r"\[(\S+), (\S+), (\S+)\]", r.stdout)
    got = np.array([[float(v) for v in row] for row in rows[-3:]])
    assert np.abs(got - ref).max() <= 1e-6
    assert sorted(os.listdir(tmp_path / "teff")) == ["effdiff_chi_dir0", "effdiff_chi_dir1", "effdiff_chi_dir2"]


def test_host_layer_on_the_mock_nan_inf_conventions(mock_env, tmp_path):
    """value()'s in-band failure conventions through the app (TortuosityHypre.cpp:764-877): a phase
    that does not percolate in the flow direction -> NaN; equal Dirichlet values -> no potential
    gradient -> +Inf; too few iterations -> not converged -> NaN."""
    import numpy as np
    from PIL import Image
    vol = np.ones((12, 14, 16), dtype=np.uint8)           # z, y, x
    vol[:, :, 7] = 0                                       # a solid wall across X
    ims = [Image.fromarray(vol[k] * 255) for k in range(vol.shape[0])]
    ims[0].save(tmp_path / "wall.tif", save_all=True, append_images=ims[1:])
    base = ["tests/inputs/diffusion_flow_through.inputs", "filename=wall.tif", f"data_path={tmp_path}/", "threshold_val=127.5",
            "verbose=0"]

    def taus(*extra, name):
        res = tmp_path / name
        run("Diffusion", *base, f"results_path={res}/", *extra, env=mock_env)
        return dict(re.findall(r"(Tortuosity_[XYZ]): (\S+)", open(res / "results.txt").read()))

    t = taus("direction=X Y", name="blocked")
    assert t["Tortuosity_X"].lower().lstrip("-") == "nan"                       # nothing percolates in X
    assert abs(float(t["Tortuosity_Y"]) - (14 - 1) / 14 * 1.0) < 1e-6           # open slabs: tau = (N-1)/N
    t = taus("direction=Y", "tortuosity.vlo=0.5", "tortuosity.vhi=0.5", name="flat")
    assert t["Tortuosity_Y"].lower() in ("inf", "nan")                          # zero gradient / zero flux
    assert t["Tortuosity_Y"].lower() == "inf"

