"""Host-side input generators of the product (sofa_b200/topology.py) against the oracle's restatement of the reference's
topology components, and the Gmsh v1 reader against the committed liver fixture."""
import os

import numpy as np
import pytest

import oracle_lib as O
from sofa_b200 import topology as T

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("n", [(4, 10, 4), (5, 5, 20), (3, 3, 4), (2, 2, 2)])
def test_grid_and_tessellations_match_reference_order(n):
    p1, h1 = T.regular_grid(n, (-5, -5, 0), (5, 5, 40)); p2, h2 = O.regular_grid(n, (-5, -5, 0), (5, 5, 40))
    assert p1.tobytes() == p2.tobytes() and np.array_equal(h1, h2)
    for mode, om in (("mapping", 0), ("mapping_swapping", 1), ("forcefield", 2)):
        assert np.array_equal(T.hexas_to_tetras(h1, n, mode), O.hexas_to_tetras(n, om)), mode


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_diagonal_mass_lumping_bit_exact(dtype):
    n = (4, 5, 9)
    pos, hexas = T.regular_grid(n, (0, 0, 0), (1, 1.3, 2.7))
    tets = T.hexas_to_tetras(hexas, n, "mapping_swapping")
    for elems in (tets, hexas):
        s = O.OracleScene(dtype, pos); s.set_mass_density(0.2, elems)
        assert T.diagonal_mass(pos, elems, dtype, mass_density=0.2).tobytes() == s.get("vertexMass").tobytes()
        s = O.OracleScene(dtype, pos); s.set_total_mass(50.0, elems)
        assert T.diagonal_mass(pos, elems, dtype, total_mass=50.0).tobytes() == s.get("vertexMass").tobytes()


def test_gmsh_v1_reader(tmp_path):
    m = np.load(os.path.join(G, "liver_mesh.npz"))
    assert m["positions"].shape == (181, 3) and m["tetrahedra"].shape == (596, 4) and m["tetrahedra"].max() == 180
    # round trip through a file written in the same $NOD/$ELM format
    p = tmp_path / "t.msh"
    with open(p, "w") as fh:
        fh.write("$NOD\n%d\n" % len(m["positions"]))
        for i, x in enumerate(m["positions"]):
            fh.write(f"{i + 1} {float(x[0])!r} {float(x[1])!r} {float(x[2])!r}\n")
        fh.write("$ENDNOD\n$ELM\n%d\n" % len(m["tetrahedra"]))
        for i, t in enumerate(m["tetrahedra"]):
            fh.write(f"{i + 1} 4 1 1 4 {t[0] + 1} {t[1] + 1} {t[2] + 1} {t[3] + 1}\n")
        fh.write("$ENDELM\n")
    pos, tets, hexas = T.read_gmsh_v1(str(p))
    assert np.array_equal(pos, m["positions"]) and np.array_equal(tets, m["tetrahedra"]) and hexas.shape == (0, 8)


def test_box_roi_closed_intervals():
    pos, _ = T.regular_grid((5, 5, 20), (-5, -5, 0), (5, 5, 40))
    idx = T.box_roi(pos, (-6, -6, -1, 50, 6, 0.1))
    assert len(idx) == 25 and np.array_equal(idx, O.box_roi(pos, (-6, -6, -1, 50, 6, 0.1)))


def test_write_state_read_state_round_trip(tmp_path):
    """WriteState / ReadState text dumps (Playback/WriteState.inl:350-381): the reference's layout, exact with precision=17."""
    rng = np.random.default_rng(4)
    frames = [dict(T=0.01 * k, X=rng.standard_normal((7, 3)), V=rng.standard_normal((7, 3))) for k in range(3)]
    p = tmp_path / "beam.state"
    T.write_state(p, frames, precision=17)
    txt = p.read_text().splitlines()
    assert txt[0].startswith("T= ") and txt[1].startswith("  X= ") and txt[2].startswith("  V= ")
    back = T.read_state(p)
    assert len(back) == 3
    for a, b in zip(frames, back):
        assert a["T"] == b["T"] and (a["X"] == b["X"]).all() and (a["V"] == b["V"]).all()
    T.write_state(p, frames)                           # the reference's default stream precision: 6 significant digits
    back = T.read_state(p)
    assert np.allclose(back[1]["X"], frames[1]["X"], rtol=1e-5, atol=1e-6)


def test_gmsh_v2_reader_matches_v1(tmp_path):
    """The same small mesh written as MSH 1.0 and as MSH 2.2 (with tags, a surface triangle and a point element to skip)."""
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 1]], float)
    tets = [[1, 2, 3, 4], [2, 3, 4, 5]]
    v1 = tmp_path / "m1.msh"
    v1.write_text("$NOD\n5\n" + "".join(f"{i + 1} {float(p[0])!r} {float(p[1])!r} {float(p[2])!r}\n" for i, p in enumerate(pos)) + "$ENDNOD\n$ELM\n2\n"
                  + "".join(f"{i + 1} 4 1 1 4 {t[0]} {t[1]} {t[2]} {t[3]}\n" for i, t in enumerate(tets)) + "$ENDELM\n")
    v2 = tmp_path / "m2.msh"
    v2.write_text("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n5\n" + "".join(f"{i + 1} {float(p[0])!r} {float(p[1])!r} {float(p[2])!r}\n" for i, p in enumerate(pos))
                  + "$EndNodes\n$Elements\n4\n1 15 2 0 1 1\n2 2 2 0 1 1 2 3\n"
                  + "".join(f"{i + 3} 4 2 0 1 {t[0]} {t[1]} {t[2]} {t[3]}\n" for i, t in enumerate(tets)) + "$EndElements\n")
    p1, t1, h1 = T.read_gmsh(v1)
    p2, t2, h2 = T.read_gmsh(v2)
    assert (p1 == p2).all() and (t1 == t2).all() and t1.tolist() == [[0, 1, 2, 3], [1, 2, 3, 4]] and h1.shape == (0, 8) and h2.shape == (0, 8)


@pytest.mark.parametrize("binary", [False, True], ids=["ascii", "binary"])
def test_vtk_legacy_unstructured_grid_reader(tmp_path, binary):
    """Legacy VTK UNSTRUCTURED_GRID, the volumetric subset of MeshVTKLoader (cell types 10 tetra / 12 hexahedron; 5 = triangle skipped)."""
    from sofa_b200 import topology as T
    pos, hexas = T.regular_grid((3, 2, 2), (0, 0, 0), (2, 1, 1))
    tets = T.hexas_to_tetras(hexas[:1], (3, 2, 2), "mapping")
    cells = [(10, t) for t in tets] + [(5, [0, 1, 2])] + [(12, h) for h in hexas[1:]]
    size = sum(1 + len(c[1]) for c in cells)
    path = tmp_path / "mesh.vtk"
    with open(path, "wb") as fh:
        fh.write(b"# vtk DataFile Version 3.0\nsofa_b200 test\n" + (b"BINARY" if binary else b"ASCII") + b"\nDATASET UNSTRUCTURED_GRID\n")
        fh.write(f"POINTS {pos.shape[0]} {'double' if binary else 'float'}\n".encode())
        if binary:
            fh.write(pos.astype(">f8").tobytes() + b"\n")
        else:
            fh.write("\n".join(" ".join(repr(float(v)) for v in p) for p in pos).encode() + b"\n")
        fh.write(f"\nCELLS {len(cells)} {size}\n".encode())
        flat = np.concatenate([np.concatenate([[len(c[1])], np.asarray(c[1], np.int64)]) for c in cells])
        fh.write(flat.astype(">i4").tobytes() + b"\n" if binary else (" ".join(str(int(v)) for v in flat) + "\n").encode())
        fh.write(f"CELL_TYPES {len(cells)}\n".encode())
        ct = np.array([c[0] for c in cells])
        fh.write(ct.astype(">i4").tobytes() + b"\n" if binary else ("\n".join(str(int(v)) for v in ct) + "\n").encode())
        fh.write(b"CELL_DATA 1\nSCALARS x float\n")
    p2, t2, h2 = T.read_vtk_legacy(str(path))
    assert np.array_equal(p2, pos) and np.array_equal(t2, tets) and np.array_equal(h2, hexas[1:])
    with open(tmp_path / "poly.vtk", "wb") as fh:
        fh.write(b"# vtk DataFile Version 3.0\nt\nASCII\nDATASET POLYDATA\nPOINTS 0 float\n")
    with pytest.raises(ValueError):
        T.read_vtk_legacy(str(tmp_path / "poly.vtk"))
