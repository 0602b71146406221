"""Host layer of libb2gpu.so (box2dpp_b200/csrc/b2g_host.cpp): the scene-time math of
b2PolygonShape::Set / SetAsBox / ComputeMass (Box2D/Collision/Shapes/b2PolygonShape.cpp:32-245, 354-437)
against the golden answers of the reference and against the oracles."""
import os

import numpy as np
import pytest

import golden_cases as G
from box2dpp_b200 import batch as B
from box2dpp_b200 import scene as S

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_polygon_build_matches_golden():
    g = np.load(os.path.join(GOLD, "polygons.npz"))
    kinds = set()
    for i, (fd, verts) in enumerate(G.polygon_cases()):
        n, out = B.host_polygon_build(fd, verts)
        kinds.add(fd.shapeKind)
        assert n == g["counts"][i], f"polygon {i}: vertex count {n} vs {g['counts'][i]}"
        assert (out.view(np.uint32) == g["outs"][i].view(np.uint32)).all(), f"polygon {i}: hull / normals / mass differ"
    assert kinds == {0, 1, 2}


def test_polygon_build_matches_oracles(oracles):
    for kind in ["port"] + (["ref"] if oracles.available("ref") else []):
        for i, (fd, verts) in enumerate(G.polygon_cases()[:60]):
            n, out = B.host_polygon_build(fd, verts)
            rn, rout = oracles.polygon_build(fd, verts, kind=kind)
            assert n == rn and (out.view(np.uint32) == rout.view(np.uint32)).all(), f"{kind} polygon {i}"


def test_fixture_mass_data_matches_reference(oracles):
    """b2gpu_host_fixture_mass_data (the payload of B2GPU_CMD_SET_DENSITY) == b2Fixture::GetMassData of the live reference after
    b2Fixture::SetDensity, bit for bit, for every fixture of scenes that hold every shape type (polygons, boxes, circles,
    edges, chains) at several densities, 0 included."""
    if not oracles.available("ref"):
        pytest.skip("oracle/_ref not built")
    seen = set()
    for scene in (S.pyramid(3), S.mixed(5, bodies=12), S.hello_world(), S.sensors(1, bodies=6), S.jointed(1), S.vehicles(2),
                  S.pyramid(3, ground="edge")):
        ref = oracles.RefWorld(scene)
        for fx, fd in enumerate(scene.fixtures):
            seen.add(int(fd.shapeType))
            for density in (float(fd.density), 0.0, 0.5, 3.25):
                ref.set_density(fx, density)
                want = ref.fixture_mass_data(fx)
                got = B.host_fixture_mass_data(fd, scene.vertices, density)
                assert want.view(np.uint32).tolist() == got.view(np.uint32).tolist(), (fx, int(fd.shapeType), density, want, got)
    assert seen == {S.SHAPE_CIRCLE, S.SHAPE_EDGE, S.SHAPE_POLYGON, S.SHAPE_CHAIN}, seen


def test_fixture_mass_data_rejects_bad_input():
    fd = S.FixtureDef()
    fd.shapeType, fd.shapeKind, fd.firstVertex, fd.vertexCount = S.SHAPE_POLYGON, S.POLY_SET, 0, 5
    with pytest.raises(B.B2GpuError):
        B.host_fixture_mass_data(fd, [0.0, 0.0, 1.0, 0.0, 0.0, 1.0])  # reads 5 vertices of 3
