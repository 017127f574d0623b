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
