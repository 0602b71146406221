"""Host layer of libb2gpu.so (box2dpp_b200/csrc/b2g_host.cpp): the scene-time math of
b2PolygonShape::Set / SetAsBox / ComputeMass (Box2D/Collision/Shapes/b2PolygonShape.cpp:32-245, 354-437)
against the golden answers of the reference and against the oracles."""
import os

import numpy as np

import golden_cases as G
from box2dpp_b200 import batch as B

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
