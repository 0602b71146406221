/* TEST INFRASTRUCTURE — CPU oracle (plain-C restatement of the reference's b2World::Step path).
 * Never linked into, imported by or executed from the product path.
 *
 * PARITY PINNED: every function here is checked bit-for-bit against the reference's own vendored
 * vanilla Box2D 2.3.2 (oracle/_ref/libb2ref.so built from /root/reference by oracle/Makefile) in
 * tests/test_oracle_vs_reference.py, and against the SURVEY Appendix D checkpoints committed in
 * tests/golden/.
 */
#ifndef B2O_H
#define B2O_H

#include "b2o_math.h"
#include "../include/b2g_scene.h"

enum { B2O_MANIFOLD_CIRCLES = 0, B2O_MANIFOLD_FACE_A = 1, B2O_MANIFOLD_FACE_B = 2 };
enum { B2O_CF_VERTEX = 0, B2O_CF_FACE = 1 };

/* b2ContactFeature / b2ContactID — Box2D/Collision/b2Collision.h:43-62 */
static inline uint32_t b2o_make_key(int indexA, int indexB, int typeA, int typeB) {
    return (uint32_t)(indexA & 0xff) | ((uint32_t)(indexB & 0xff) << 8) | ((uint32_t)(typeA & 0xff) << 16) |
           ((uint32_t)(typeB & 0xff) << 24);
}

/* b2ManifoldPoint / b2Manifold — Box2D/Collision/b2Collision.h:74-112 */
typedef struct {
    v2 localPoint;
    float normalImpulse, tangentImpulse;
    uint32_t key;
} b2o_mpoint;

typedef struct {
    b2o_mpoint points[2];
    v2 localNormal, localPoint;
    int type;
    int pointCount;
} b2o_manifold;

/* Shapes after construction (b2PolygonShape.h:88-96, b2CircleShape.h, b2EdgeShape.h, b2ChainShape.h) */
typedef struct {
    int type; /* B2G_SHAPE_* */
    float radius;
    /* polygon */
    int count;
    v2 verts[B2O_MAX_POLY], normals[B2O_MAX_POLY], centroid;
    /* circle */
    v2 p;
    /* edge (also a chain's child, b2ChainShape::GetChildEdge) */
    v2 v0, v1, v2_, v3;
    int has0, has3;
    /* chain */
    v2* chain;
    int chainCount;
    v2 prev, next;
    int hasPrev, hasNext;
} b2o_shape;

typedef struct { float mass; v2 center; float I; } b2o_massdata;

void b2o_polygon_set_box(b2o_shape* s, float hx, float hy);
void b2o_polygon_set_box_offset(b2o_shape* s, float hx, float hy, v2 center, float angle);
void b2o_polygon_set(b2o_shape* s, const v2* pts, int count);
void b2o_shape_mass(const b2o_shape* s, float density, b2o_massdata* md);
void b2o_shape_aabb(const b2o_shape* s, xform xf, int child, aabb* out);
void b2o_chain_child_edge(const b2o_shape* chain, int index, b2o_shape* edge);

void b2o_collide_polygons_m(b2o_manifold* m, const b2o_shape* a, xform xfA, const b2o_shape* b, xform xfB);
void b2o_collide_circles_m(b2o_manifold* m, const b2o_shape* a, xform xfA, const b2o_shape* b, xform xfB);
void b2o_collide_polygon_circle_m(b2o_manifold* m, const b2o_shape* a, xform xfA, const b2o_shape* b, xform xfB);
void b2o_collide_edge_circle_m(b2o_manifold* m, const b2o_shape* a, xform xfA, const b2o_shape* b, xform xfB);
void b2o_collide_edge_polygon_m(b2o_manifold* m, const b2o_shape* a, xform xfA, const b2o_shape* b, xform xfB);

#endif
