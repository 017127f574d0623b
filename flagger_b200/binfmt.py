"""Readers/writers for the file formats on either side of the hot path.

* `.bin` chunk dump: ChunksCreator_writeChunksIntoBinaryFile / ..._parseChunksFromBinaryFile
  (submodules/chunk/chunk.c:596-709, 713-828).  Little-endian, C `bool` = 1 byte.
* `.cov` / `.cov.gz`: header keys per submodules/track_reader/track_reader.c:220-457; data lines
  `start end cov cov_high_mapq cov_high_clip annotIdx[,..] regionIdx [truth [prediction]]`, 1-based inclusive.
* alpha TSV: 4x4, no header, every line newline-terminated (submodules/data_types/data_types.c:490-518).
"""
import gzip
import struct

import numpy as np

from . import _abi
from .synth import Workload

REGION_SHIFT = 58  # ptBlock.c:294-304: the top 6 bits of annotation_flag hold the region index


def annotation_flags(wl):
    """u64 flag per window: bit (k-1) for annotation k>=1; region index in bits 58..63 (ptBlock.c:225-228,294-304)."""
    flag = np.full(wl.n_windows, 1 << 0, dtype=np.uint64)  # annotation 1 = whole_genome
    if wl.n_regions > 1:
        flag |= (np.uint64(1) << (wl.region.astype(np.uint64) + np.uint64(1))) * (wl.region > 0).astype(np.uint64)
    flag |= wl.region.astype(np.uint64) << np.uint64(REGION_SHIFT)
    return flag


def write_bin(wl, path, with_truth=False):
    with open(path, "wb") as f:
        f.write(struct.pack("<i", len(wl.annotation_names)))
        for name in wl.annotation_names:
            b = name.encode() + b"\0"
            f.write(struct.pack("<i", len(b)))
            f.write(b)
        f.write(struct.pack("<i", wl.n_regions))
        f.write(wl.region_coverages.astype("<i4").tobytes())
        f.write(struct.pack("<i", 4 if with_truth else 0))       # numberOfLabels
        f.write(struct.pack("<???", bool(with_truth), False, False))  # truth, prediction, startOnly
        f.write(struct.pack("<i", wl.avg_alignment_len))
        f.write(struct.pack("<ii", wl.chunk_len, wl.window_len))
        flags = annotation_flags(wl)
        for c, name in zip(wl.chunks, wl.contig_names):
            b = name.encode() + b"\0"
            o, n = int(c["offset"]), int(c["n_windows"])
            f.write(struct.pack("<i", len(b)))
            f.write(b)
            f.write(struct.pack("<iiii", int(c["ctg_len"]), int(c["s"]), int(c["e"]), n))
            f.write(wl.cov[o:o + n].astype("<u2").tobytes())
            f.write(wl.cov_high_mapq[o:o + n].astype("<u2").tobytes())
            f.write(wl.cov_high_clip[o:o + n].astype("<u2").tobytes())
            f.write(flags[o:o + n].astype("<u8").tobytes())
            truth = wl.truth[o:o + n] if with_truth else np.full(n, -1, np.int8)
            f.write(truth.astype("i1").tobytes())
            f.write(np.full(n, -1, np.int8).tobytes())


def read_bin(path):
    """Parse a `.bin` chunk dump into a Workload (+ header dict)."""
    data = open(path, "rb").read()
    p = 0

    def take(fmt):
        nonlocal p
        v = struct.unpack_from(fmt, data, p)
        p += struct.calcsize(fmt)
        return v

    (n_ann,) = take("<i")
    ann = []
    for _ in range(n_ann):
        (ln,) = take("<i")
        ann.append(data[p:p + ln - 1].decode())
        p += ln
    (n_reg,) = take("<i")
    reg_cov = np.frombuffer(data, "<i4", n_reg, p).copy()
    p += 4 * n_reg
    (n_labels,) = take("<i")
    truth_avail, pred_avail, start_only = take("<???")
    (avg_len,) = take("<i")
    chunk_len, window_len = take("<ii")
    descs, names, cov, mq, cl, reg, tr = [], [], [], [], [], [], []
    off = 0
    while p < len(data):
        (ln,) = take("<i")
        names.append(data[p:p + ln - 1].decode())
        p += ln
        ctg_len, s, e, n = take("<iiii")
        cov.append(np.frombuffer(data, "<u2", n, p)); p += 2 * n
        mq.append(np.frombuffer(data, "<u2", n, p)); p += 2 * n
        cl.append(np.frombuffer(data, "<u2", n, p)); p += 2 * n
        fl = np.frombuffer(data, "<u8", n, p); p += 8 * n
        reg.append((fl >> np.uint64(REGION_SHIFT)).astype(np.uint8))
        tr.append(np.frombuffer(data, "i1", n, p)); p += n
        p += n  # prediction
        descs.append((ctg_len, s, e, window_len, n, 0, off))
        off += n
    chunks = np.array(descs, dtype=_abi.chunk_desc_dtype)
    cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)
    wl = Workload(path, window_len, chunk_len, avg_len, reg_cov, names, chunks, cat(cov, np.uint16), cat(mq, np.uint16),
                  cat(cl, np.uint16), cat(reg, np.uint8), cat(tr, np.int8), ann)
    hdr = dict(n_labels=n_labels, truth=truth_avail, prediction=pred_avail, start_only=start_only)
    return wl, hdr


def write_cov(wl, path, with_truth=False):
    """Run-length `.cov`/`.cov.gz`: one block per window, so the window mean equals the window value."""
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "wt") as f:
        f.write(f"#annotation:len:{len(wl.annotation_names)}\n")
        for i, name in enumerate(wl.annotation_names):
            f.write(f"#annotation:name:{i}:{name}\n")
        f.write(f"#region:len:{wl.n_regions}\n")
        for i, c in enumerate(wl.region_coverages):
            f.write(f"#region:coverage:{i}:{int(c)}\n")
        if with_truth:
            f.write("#label:len:4\n")
            for i, n in enumerate(_abi.STATE_NAMES):
                f.write(f"#label:name:{i}:{n}\n")
        else:
            f.write("#label:len:0\n")
        f.write(f"#truth:{'true' if with_truth else 'false'}\n#prediction:false\n")
        f.write(f"#avg_alignment_len:{wl.avg_alignment_len}\n#start-only:false\n")
        prev = None
        for c, name in zip(wl.chunks, wl.contig_names):
            if name != prev:
                f.write(f">{name} {int(c['ctg_len'])}\n")
                prev = name
            o, n, s, e, w = int(c["offset"]), int(c["n_windows"]), int(c["s"]), int(c["e"]), int(c["window_len"])
            for i in range(n):
                a = s + i * w
                b = min(a + w - 1, e)
                r = int(wl.region[o + i])
                annot = "1" if r == 0 else f"1,{r + 1}"
                line = f"{a + 1}\t{b + 1}\t{int(wl.cov[o + i])}\t{int(wl.cov_high_mapq[o + i])}\t" \
                       f"{int(wl.cov_high_clip[o + i])}\t{annot}\t{r}"
                if with_truth:
                    line += f"\t{int(wl.truth[o + i])}"
                f.write(line + "\n")


def write_alpha_tsv(alpha, path):
    with open(path, "w") as f:
        for row in np.asarray(alpha).reshape(4, 4):
            f.write("\t".join(f"{v:.3f}" for v in row) + "\n")


# ---- native readers (flagger_b200/csrc/hfg_cov_reader.c through include/hfg_io.h) --------------------------------------

import ctypes as _C


class _CovData(_C.Structure):
    _fields_ = [
        ("n_annotations", _C.c_int32), ("annotation_names", _C.POINTER(_C.c_char_p)),
        ("n_regions", _C.c_int32), ("region_coverages", _C.POINTER(_C.c_int32)),
        ("n_labels", _C.c_int32), ("truth_available", _C.c_int32), ("prediction_available", _C.c_int32),
        ("start_only", _C.c_int32), ("avg_alignment_len", _C.c_int32),
        ("chunk_len", _C.c_int32), ("window_len", _C.c_int32),
        ("n_chunks", _C.c_int32), ("chunks", _C.c_void_p), ("contig_names", _C.c_void_p),
        ("n_windows", _C.c_int64),
        ("cov", _C.POINTER(_C.c_uint16)), ("cov_high_mapq", _C.POINTER(_C.c_uint16)),
        ("cov_high_clip", _C.POINTER(_C.c_uint16)), ("annotation_flag", _C.POINTER(_C.c_uint64)),
        ("region", _C.POINTER(_C.c_uint8)), ("truth", _C.POINTER(_C.c_int8)), ("prediction", _C.POINTER(_C.c_int8)),
    ]


def _to_workload(lib, dptr, name):
    d = dptr.contents
    W, Cn = int(d.n_windows), int(d.n_chunks)
    arr = lambda p, dt: np.ctypeslib.as_array(p, shape=(W,)).astype(dt, copy=True) if W else np.zeros(0, dt)
    chunks = np.frombuffer(_C.string_at(d.chunks, Cn * _abi.chunk_desc_dtype.itemsize), dtype=_abi.chunk_desc_dtype).copy()
    raw = _C.string_at(d.contig_names, Cn * 200)
    names = [raw[i * 200:(i + 1) * 200].split(b"\0", 1)[0].decode() for i in range(Cn)]
    wl = Workload(name, int(d.window_len), int(d.chunk_len), int(d.avg_alignment_len),
                  np.array([d.region_coverages[i] for i in range(d.n_regions)], np.int32), names, chunks,
                  arr(d.cov, np.uint16), arr(d.cov_high_mapq, np.uint16), arr(d.cov_high_clip, np.uint16),
                  arr(d.region, np.uint8), arr(d.truth, np.int8),
                  [d.annotation_names[i].decode() for i in range(d.n_annotations)])
    hdr = dict(n_labels=int(d.n_labels), truth=bool(d.truth_available), prediction=bool(d.prediction_available),
               start_only=bool(d.start_only), annotation_flag=arr(d.annotation_flag, np.uint64),
               prediction_labels=arr(d.prediction, np.int8))
    lib.hfg_cov_free(dptr)
    return wl, hdr


def _io_lib():
    from .api import lib
    L = lib()
    L.hfg_cov_free.restype = None
    L.hfg_cov_free.argtypes = [_C.POINTER(_CovData)]
    return L


def read_cov_native(path, chunk_len=20_000_000, window_len=4000):
    """`.cov` / `.cov.gz` -> Workload via the C reader (hfg_read_cov): one pass, run-length aware."""
    L = _io_lib()
    out = _C.POINTER(_CovData)()
    err = _C.create_string_buffer(512)
    rc = L.hfg_read_cov(str(path).encode(), _C.c_int32(chunk_len), _C.c_int32(window_len), _C.byref(out), err, _C.c_size_t(512))
    if rc != 0:
        raise ValueError(f"hfg_read_cov: {err.value.decode()}")
    return _to_workload(L, out, str(path))


def read_bin_native(path):
    """`.bin` chunk dump -> Workload via the C reader (hfg_read_bin)."""
    L = _io_lib()
    out = _C.POINTER(_CovData)()
    err = _C.create_string_buffer(512)
    rc = L.hfg_read_bin(str(path).encode(), _C.byref(out), err, _C.c_size_t(512))
    if rc != 0:
        raise ValueError(f"hfg_read_bin: {err.value.decode()}")
    return _to_workload(L, out, str(path))


def write_random_rle_cov(path, contig_lens, seed=0, n_regions=3, with_truth=True, float_values=False,
                         avg_alignment_len=15000):
    """A `.cov`/`.cov.gz` whose run-length blocks do NOT line up with windows or chunks (lengths 1..3000), with several
    annotations per block, regions and truth labels: exercises the window builder's averaging / mode / OR logic."""
    rng = np.random.default_rng(seed)
    opener = gzip.open if str(path).endswith(".gz") else open
    region_cov = [40, 52, 30, 61, 25, 48, 36][:n_regions]
    with opener(path, "wt") as f:
        f.write("#annotation:len:5\n")
        for i, nm in enumerate(["no_annotation", "whole_genome", "sat_a", "sat_b", "sat_c"]):
            f.write(f"#annotation:name:{i}:{nm}\n")
        f.write(f"#region:len:{n_regions}\n")
        for i, c in enumerate(region_cov):
            f.write(f"#region:coverage:{i}:{c}\n")
        f.write(f"#label:len:{4 if with_truth else 0}\n")
        if with_truth:
            for i, nm in enumerate(_abi.STATE_NAMES):
                f.write(f"#label:name:{i}:{nm}\n")
        f.write(f"#truth:{'true' if with_truth else 'false'}\n#prediction:false\n")
        f.write(f"#avg_alignment_len:{avg_alignment_len}\n#start-only:false\n")
        for ci, L in enumerate(contig_lens):
            f.write(f">ctg{ci + 1} {int(L)}\n")
            pos = 1
            while pos <= L:
                ln = int(min(rng.integers(1, 3001), L - pos + 1))
                cov = float(rng.integers(0, 300))
                if float_values:
                    cov += float(rng.integers(0, 4)) / 4 + (0.1 if rng.random() < 0.3 else 0.0)
                mq = cov * float(rng.choice([0.0, 0.1, 0.5, 1.0]))
                cl = cov * float(rng.choice([0.0, 0.0, 1.0]))
                fmt = (lambda v: f"{v:.2f}") if float_values else (lambda v: f"{int(v)}")
                r = int(rng.integers(0, n_regions))
                annots = sorted(set([1] + list(rng.integers(1, 5, size=int(rng.integers(0, 3))))))
                line = f"{pos}\t{pos + ln - 1}\t{fmt(cov)}\t{fmt(mq)}\t{fmt(cl)}\t{','.join(str(a) for a in annots)}\t{r}"
                if with_truth:
                    line += f"\t{int(rng.integers(-1, 4))}"
                f.write(line + "\n")
                pos += ln


def write_summary_native(inp, out_path, prediction=None, use_truth=True, label_names=_abi.STATE_NAMES + ("Unk",),
                         overlap_ratio_threshold=0.4, chunk_len=20_000_000, window_len=4000, bin_array_file=None):
    """prediction_summary_<suffix>.tsv through the C writer (hfg_write_summary_tsv, csrc/hfg_summary.c) for the input file
    `inp` (.cov / .cov.gz / .bin), the per-window `prediction` labels (int8, or None) and the file's own truth labels."""
    L = _io_lib()
    out = _C.POINTER(_CovData)()
    err = _C.create_string_buffer(512)
    if str(inp).endswith(".bin"):
        rc = L.hfg_read_bin(str(inp).encode(), _C.byref(out), err, _C.c_size_t(512))
    else:
        rc = L.hfg_read_cov(str(inp).encode(), _C.c_int32(chunk_len), _C.c_int32(window_len), _C.byref(out), err, _C.c_size_t(512))
    if rc != 0:
        raise ValueError(f"reader: {err.value.decode()}")
    try:
        d = out.contents
        truth = d.truth if (use_truth and d.truth_available) else None
        pred = None
        if prediction is not None:
            pred = np.ascontiguousarray(prediction, np.int8)
            assert pred.shape[0] == int(d.n_windows)
        names = None
        n_labels = len(_abi.STATE_NAMES)
        if label_names is not None:
            names = (_C.c_char_p * len(label_names))(*[s.encode() for s in label_names])
            n_labels = len(label_names) - 1
        rc = L.hfg_write_summary_tsv(str(out_path).encode(), out, _abi.ptr(pred), truth, names, _C.c_int(n_labels),
                                     _C.c_double(overlap_ratio_threshold),
                                     str(bin_array_file).encode() if bin_array_file else None, err, _C.c_size_t(512))
        if rc != 0:
            raise ValueError(f"hfg_write_summary_tsv: {err.value.decode()}")
    finally:
        L.hfg_cov_free(out)


class NativeCov:
    """An input file parsed ONCE by the C reader and kept alive: `.workload` (numpy view of the windows, what
    HmmFlaggerGPU takes), `.truth_available`, and the writers / scorers that need the parsed structure
    (`write_summary`, `benchmark_scores`).  Call `.close()` (or let it be collected) to free the C side."""

    def __init__(self, path, chunk_len=20_000_000, window_len=4000):
        self._L = _io_lib()
        self._p = _C.POINTER(_CovData)()
        err = _C.create_string_buffer(512)
        if str(path).endswith(".bin"):
            rc = self._L.hfg_read_bin(str(path).encode(), _C.byref(self._p), err, _C.c_size_t(512))
        else:
            rc = self._L.hfg_read_cov(str(path).encode(), _C.c_int32(chunk_len), _C.c_int32(window_len), _C.byref(self._p), err,
                                      _C.c_size_t(512))
        if rc != 0:
            raise ValueError(f"reader: {err.value.decode()}")
        d = self._p.contents
        self.path = str(path)
        self.n_windows = int(d.n_windows)
        self.truth_available = bool(d.truth_available)
        W, Cn = self.n_windows, int(d.n_chunks)
        arr = lambda p, dt: np.ctypeslib.as_array(p, shape=(W,)).astype(dt, copy=True) if W else np.zeros(0, dt)
        chunks = np.frombuffer(_C.string_at(d.chunks, Cn * _abi.chunk_desc_dtype.itemsize), dtype=_abi.chunk_desc_dtype).copy()
        raw = _C.string_at(d.contig_names, Cn * 200)
        names = [raw[i * 200:(i + 1) * 200].split(b"\0", 1)[0].decode() for i in range(Cn)]
        self.workload = Workload(self.path, int(d.window_len), int(d.chunk_len), int(d.avg_alignment_len),
                                 np.array([d.region_coverages[i] for i in range(d.n_regions)], np.int32), names, chunks,
                                 arr(d.cov, np.uint16), arr(d.cov_high_mapq, np.uint16), arr(d.cov_high_clip, np.uint16),
                                 arr(d.region, np.uint8), arr(d.truth, np.int8),
                                 [d.annotation_names[i].decode() for i in range(d.n_annotations)])

    def close(self):
        if self._p:
            self._L.hfg_cov_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def benchmark_scores(self, prediction, annotation_label="whole_genome", size_label="ALL_SIZES",
                         overlap_ratio_threshold=0.4, bin_array_file=None):
        """(overlap-based F1, base-level F1, contiguity) of `prediction` against the file's truth labels: the three numbers
        the reference's alpha-tuning driver reads from the benchmarking files of a run (hfg_benchmark_scores)."""
        if not self.truth_available:
            raise ValueError(f"{self.path} carries no truth labels")
        pred = np.ascontiguousarray(prediction, np.int8)
        assert pred.shape[0] == self.n_windows
        scores = (_C.c_double * 3)()
        err = _C.create_string_buffer(512)
        rc = self._L.hfg_benchmark_scores(self._p, _abi.ptr(pred), self._p.contents.truth, _C.c_int(len(_abi.STATE_NAMES)),
                                          _C.c_double(overlap_ratio_threshold),
                                          str(bin_array_file).encode() if bin_array_file else None, annotation_label.encode(),
                                          size_label.encode(), scores, err, _C.c_size_t(512))
        if rc != 0:
            raise ValueError(f"hfg_benchmark_scores: {err.value.decode()}")
        return tuple(float(v) for v in scores)
