"""B2TRK track files: the flattened SoA image of an OpenMOC TrackGenerator.

The file is what ``openmoc_b200/cpp/b200_flatten.cpp:b200_write_trackfile``
writes on the OpenMOC side of the boundary and what ``B200Solver`` uploads to the
device through the C-ABI (``include/b200moc.h``).  It plays the role of the
reference's (currently disabled) segment dump, ``TrackGenerator::dumpSegmentsToFile``
(``src/TrackGenerator.cpp:1388-1645``), but in a GPU-ready structure-of-arrays
form: one contiguous stream per field instead of one record per segment.

Layout: magic ``B2TRK001``; int64 chunk count; per chunk ``char name[24]``,
``char dtype[8]`` (numpy-style: f8, f4, i8, i4, u1), int64 element count, raw
little-endian data padded to 8 bytes.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Dict

import numpy as np

MAGIC = b"B2TRK001"

SCALARS = ("num_groups", "num_azim", "num_polar", "solve_3d", "fluxes_per_track",
           "n_tracks", "n_segments", "n_fsrs", "n_materials")

#: boundaryType enum of the reference (src/boundary_type.h:14-29)
VACUUM, REFLECTIVE, PERIODIC, INTERFACE = 0, 1, 2, 3


@dataclass
class FlatTracks:
    """Host SoA arrays; attribute names equal the chunk names of the file."""
    num_groups: int = 0
    num_azim: int = 0
    num_polar: int = 0
    solve_3d: int = 0
    fluxes_per_track: int = 0
    n_tracks: int = 0
    n_segments: int = 0
    n_fsrs: int = 0
    n_materials: int = 0
    arrays: Dict[str, np.ndarray] = field(default_factory=dict)

    def __getattr__(self, name):
        arrays = self.__dict__.get("arrays", {})
        if name in arrays:
            return arrays[name]
        raise AttributeError(name)

    @property
    def num_polar_per_track(self) -> int:
        """P/2 in 2D (polar angles carried per track), 1 in 3D."""
        return 1 if self.solve_3d else self.num_polar // 2

    def validate(self) -> None:
        a = self.arrays
        nt, ns, nf, nm, G = self.n_tracks, self.n_segments, self.n_fsrs, self.n_materials, self.num_groups
        assert self.fluxes_per_track == G * self.num_polar_per_track
        assert a["seg_length"].shape == (ns,) and a["seg_fsr"].shape == (ns,)
        assert a["trk_seg_offset"].shape == (nt + 1,)
        assert a["trk_seg_offset"][0] == 0 and a["trk_seg_offset"][-1] == ns
        assert np.all(np.diff(a["trk_seg_offset"]) >= 0)
        for k in ("trk_azim", "trk_polar", "trk_next_fwd", "trk_next_bwd", "trk_flags",
                  "trk_bc_fwd", "trk_bc_bwd"):
            assert a[k].shape == (nt,), k
        if ns:
            assert a["seg_fsr"].min() >= 0 and a["seg_fsr"].max() < nf
        assert a["fsr_volume"].shape == (nf,) and a["fsr_mat"].shape == (nf,)
        assert a["quad_weight"].shape == (self.num_azim // 2 * self.num_polar,)
        assert a["mat_sigma_t"].shape == (nm * G,)
        assert a["mat_sigma_s"].shape == (nm * G * G,)
        linked = (a["trk_bc_fwd"] == REFLECTIVE) | (a["trk_bc_fwd"] == PERIODIC)
        if linked.any():
            nx = a["trk_next_fwd"][linked]
            assert nx.min() >= 0 and nx.max() < nt
        linked = (a["trk_bc_bwd"] == REFLECTIVE) | (a["trk_bc_bwd"] == PERIODIC)
        if linked.any():
            nx = a["trk_next_bwd"][linked]
            assert nx.min() >= 0 and nx.max() < nt


def read_trackfile(path: str) -> FlatTracks:
    with open(path, "rb") as f:
        buf = f.read()
    if buf[:8] != MAGIC:
        raise ValueError(f"{path}: not a B2TRK001 track file")
    (n_chunks,) = struct.unpack_from("<q", buf, 8)
    off = 16
    ft = FlatTracks()
    for _ in range(n_chunks):
        name = buf[off:off + 24].split(b"\0", 1)[0].decode()
        dtype = buf[off + 24:off + 32].split(b"\0", 1)[0].decode()
        (count,) = struct.unpack_from("<q", buf, off + 32)
        off += 40
        dt = np.dtype("<" + dtype)
        nbytes = count * dt.itemsize
        arr = np.frombuffer(buf, dtype=dt, count=count, offset=off).copy()
        off += nbytes + (-nbytes) % 8
        if name in SCALARS:
            setattr(ft, name, int(arr[0]))
        else:
            ft.arrays[name] = arr
    return ft


def write_trackfile(ft: FlatTracks, path: str) -> None:
    chunks = [(k, np.array([getattr(ft, k)], dtype="<i8")) for k in SCALARS]
    chunks += list(ft.arrays.items())
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<q", len(chunks)))
        for name, arr in chunks:
            arr = np.ascontiguousarray(arr)
            code = arr.dtype.str.lstrip("<|=")
            f.write(name.encode().ljust(24, b"\0")[:24])
            f.write(code.encode().ljust(8, b"\0")[:8])
            f.write(struct.pack("<q", arr.size))
            raw = arr.tobytes()
            f.write(raw)
            f.write(b"\0" * ((-len(raw)) % 8))
