/* TEST INFRASTRUCTURE — CPU oracle: world construction and b2World::Step, restated in plain C over
 * flat arrays.  See b2o.h for the parity statement.  Never linked into or called from the product.
 *
 * Paths cited are relative to /root/reference/Regression/box2d-ref/Box2DRef (vanilla; bit-identical
 * to the fork's Box2D/ for this path).
 *
 * How lists are represented: the reference head-inserts every new contact into
 * the world list and into both bodies' edge lists (Dynamics/b2ContactManager.cpp:259-291) and unlinks in
 * place (:42-105).  So "list order" is always "newest first" and a body's edge list is the sub-sequence
 * of the world list that touches that body.  Here the contacts live in one array in creation order;
 * list order is that array walked backwards, and removal is a stable erase.
 *
 * The broad phase keeps no tree: SURVEY.md F9 — UpdatePairs returns every leaf whose fat AABB overlaps
 * the query's, then sorts by (proxyIdA, proxyIdB), so only fat AABBs, proxy ids and the move buffer matter.
 */
#include <stdlib.h>
#include <string.h>

#include "b2o.h"

enum { BF_ISLAND = 1, BF_AWAKE = 2, BF_AUTOSLEEP = 4, BF_BULLET = 8, BF_FIXEDROT = 16, BF_ACTIVE = 32 };
enum { CF_ISLAND = 1, CF_TOUCHING = 2, CF_ENABLED = 4, CF_FILTER = 8 };

typedef struct {
    int type, flags;
    xform xf;
    v2 localCenter, c0, c;
    float a0, a;
    v2 v;
    float w;
    v2 force;
    float torque;
    float mass, invMass, I, invI;
    float linearDamping, angularDamping, gravityScale, sleepTime;
    int islandIndex;
} body_t;

typedef struct {
    int body;
    b2o_shape shape;
    float friction, restitution, density;
    int sensor;
    uint16_t cat, mask;
    int16_t group;
    int firstProxy, proxyCount;
} fixture_t;

typedef struct {
    int fixture, child, id;
    aabb fat;
} proxy_t;

typedef struct {
    int fixtureA, fixtureB, childA, childB;
    int proxyA, proxyB;
    int flags;
    b2o_manifold m;
    float friction, restitution, tangentSpeed;
} contact_t;

typedef struct {
    int32_t type, fixtureA, childA, fixtureB, childB, count;
    float normalImpulses[2], tangentImpulses[2];
} event_t;

typedef struct {
    v2 gravity;
    int flags;
    int newFixture;
    float inv_dt0;
    int nb, nf, np, nc, capc;
    body_t* bodies;
    fixture_t* fixtures;
    proxy_t* proxies;
    contact_t* contacts; /* creation order; list order = reverse */
    v2* chainPool;
    int* moveBuffer;
    int moveCount, moveCap;
    event_t* events;
    int eventCount, eventCap, recordEvents;
    /* last-step diagnostics */
    int* islandBodies;   /* concatenated islands, body indices */
    int* islandContacts; /* concatenated islands, contact array indices at solve time */
    int islandBodyCount, islandContactCount, islandCount;
} world_t;

static void push_event(world_t* w, int type, const contact_t* c, int count, const float* ni, const float* ti) {
    if (!w->recordEvents) return;
    if (w->eventCount == w->eventCap) {
        w->eventCap = w->eventCap ? 2 * w->eventCap : 1024;
        w->events = (event_t*)realloc(w->events, sizeof(event_t) * w->eventCap);
    }
    event_t* e = &w->events[w->eventCount++];
    memset(e, 0, sizeof(*e));
    e->type = type;
    e->fixtureA = c->fixtureA; e->childA = c->childA;
    e->fixtureB = c->fixtureB; e->childB = c->childB;
    e->count = count;
    for (int i = 0; i < count && ni; ++i) { e->normalImpulses[i] = ni[i]; e->tangentImpulses[i] = ti[i]; }
}

/* b2Body::SetAwake — Dynamics/b2Body.h:634-653 */
static void set_awake(body_t* b, int flag) {
    if (flag) {
        if ((b->flags & BF_AWAKE) == 0) {
            b->flags |= BF_AWAKE;
            b->sleepTime = 0.0f;
        }
    } else {
        b->flags &= ~BF_AWAKE;
        b->sleepTime = 0.0f;
        b->v = V2(0.0f, 0.0f);
        b->w = 0.0f;
        b->force = V2(0.0f, 0.0f);
        b->torque = 0.0f;
    }
}

static void buffer_move(world_t* w, int proxy) { /* b2BroadPhase::BufferMove — Collision/b2BroadPhase.cpp:71-84 */
    if (w->moveCount == w->moveCap) {
        w->moveCap = w->moveCap ? 2 * w->moveCap : 16;
        w->moveBuffer = (int*)realloc(w->moveBuffer, sizeof(int) * w->moveCap);
    }
    w->moveBuffer[w->moveCount++] = proxy;
}

/* b2Body::ResetMassData — Dynamics/b2Body.cpp:283-352 */
static void reset_mass_data(world_t* w, int bi) {
    body_t* b = &w->bodies[bi];
    b->mass = 0.0f; b->invMass = 0.0f; b->I = 0.0f; b->invI = 0.0f;
    b->localCenter = V2(0.0f, 0.0f);
    if (b->type == B2G_STATIC_BODY || b->type == B2G_KINEMATIC_BODY) {
        b->c0 = b->xf.p;
        b->c = b->xf.p;
        b->a0 = b->a;
        return;
    }
    v2 localCenter = V2(0.0f, 0.0f);
    for (int f = w->nf - 1; f >= 0; --f) { /* fixture list is head-inserted: newest first */
        fixture_t* fx = &w->fixtures[f];
        if (fx->body != bi) continue;
        if (fx->density == 0.0f) continue;
        b2o_massdata md;
        b2o_shape_mass(&fx->shape, fx->density, &md);
        b->mass += md.mass;
        localCenter = v2_add(localCenter, v2_scale(md.mass, md.center));
        b->I += md.I;
    }
    if (b->mass > 0.0f) {
        b->invMass = 1.0f / b->mass;
        localCenter = v2_scale(b->invMass, localCenter);
    } else {
        b->mass = 1.0f;
        b->invMass = 1.0f;
    }
    if (b->I > 0.0f && (b->flags & BF_FIXEDROT) == 0) {
        b->I -= b->mass * v2_dot(localCenter, localCenter);
        b->invI = 1.0f / b->I;
    } else {
        b->I = 0.0f;
        b->invI = 0.0f;
    }
    v2 oldCenter = b->c;
    b->localCenter = localCenter;
    b->c0 = b->c = xf_mul(b->xf, b->localCenter);
    b->v = v2_add(b->v, v2_cross_sv(b->w, v2_sub(b->c, oldCenter)));
}

void* b2o_world_create(const b2g_scene* s) {
    if (s->jointCount > 0) return NULL; /* joints are checked against the live reference only */
    world_t* w = (world_t*)calloc(1, sizeof(world_t));
    w->gravity = V2(s->gravity[0], s->gravity[1]);
    w->flags = s->worldFlags;
    w->nb = s->bodyCount;
    w->bodies = (body_t*)calloc(s->bodyCount > 0 ? s->bodyCount : 1, sizeof(body_t));
    w->fixtures = (fixture_t*)calloc(s->fixtureCount > 0 ? s->fixtureCount : 1, sizeof(fixture_t));
    w->chainPool = (v2*)calloc((s->vertexCount + 1) * 2, sizeof(v2));
    int chainUsed = 0;
    int proxyCap = 0;
    for (int i = 0; i < s->fixtureCount; ++i) {
        const b2g_fixture_def* d = &s->fixtures[i];
        if (d->shapeType == B2G_SHAPE_CHAIN) proxyCap += d->vertexCount; /* loop: count, open: count-1 */
        else proxyCap += 1;
    }
    w->proxies = (proxy_t*)calloc(proxyCap > 0 ? proxyCap : 1, sizeof(proxy_t));

    /* b2Body::b2Body — Dynamics/b2Body.cpp:27-108 */
    for (int i = 0; i < s->bodyCount; ++i) {
        const b2g_body_def* d = &s->bodies[i];
        body_t* b = &w->bodies[i];
        b->type = d->type;
        b->flags = 0;
        if (d->flags & B2G_BD_BULLET) b->flags |= BF_BULLET;
        if (d->flags & B2G_BD_FIXED_ROTATION) b->flags |= BF_FIXEDROT;
        if (d->flags & B2G_BD_ALLOW_SLEEP) b->flags |= BF_AUTOSLEEP;
        if (d->flags & B2G_BD_AWAKE) b->flags |= BF_AWAKE;
        if (d->flags & B2G_BD_ACTIVE) b->flags |= BF_ACTIVE;
        b->xf.p = V2(d->position[0], d->position[1]);
        b->xf.q = rot_set(d->angle);
        b->localCenter = V2(0.0f, 0.0f);
        b->c0 = b->xf.p;
        b->c = b->xf.p;
        b->a0 = d->angle;
        b->a = d->angle;
        b->v = V2(d->linearVelocity[0], d->linearVelocity[1]);
        b->w = d->angularVelocity;
        b->linearDamping = d->linearDamping;
        b->angularDamping = d->angularDamping;
        b->gravityScale = d->gravityScale;
        if (b->type == B2G_DYNAMIC_BODY) { b->mass = 1.0f; b->invMass = 1.0f; }
    }

    /* b2Body::CreateFixture — Dynamics/b2Body.cpp:168-204; b2Fixture::Create / CreateProxies —
     * Dynamics/b2Fixture.cpp:43-69, 124-139; b2DynamicTree::CreateProxy — Collision/b2DynamicTree.cpp:108-123 */
    for (int i = 0; i < s->fixtureCount; ++i) {
        const b2g_fixture_def* d = &s->fixtures[i];
        fixture_t* f = &w->fixtures[i];
        w->nf = i + 1;
        f->body = d->body;
        f->friction = d->friction;
        f->restitution = d->restitution;
        f->density = d->density;
        f->sensor = d->isSensor;
        f->cat = d->categoryBits; f->mask = d->maskBits; f->group = d->groupIndex;
        b2o_shape* sh = &f->shape;
        const float* V = s->vertices;
        int childCount = 1;
        switch (d->shapeType) {
            case B2G_SHAPE_CIRCLE:
                sh->type = B2G_SHAPE_CIRCLE;
                sh->radius = d->radius;
                sh->p = V2(d->p[0], d->p[1]);
                break;
            case B2G_SHAPE_EDGE:
                sh->type = B2G_SHAPE_EDGE;
                sh->radius = B2O_POLYGON_RADIUS;
                sh->v1 = V2(V[2 * d->firstVertex], V[2 * d->firstVertex + 1]);
                sh->v2_ = V2(V[2 * d->firstVertex + 2], V[2 * d->firstVertex + 3]);
                sh->has0 = d->hasVertex0; sh->has3 = d->hasVertex3;
                sh->v0 = V2(d->vertex0[0], d->vertex0[1]);
                sh->v3 = V2(d->vertex3[0], d->vertex3[1]);
                break;
            case B2G_SHAPE_POLYGON:
                if (d->shapeKind == B2G_POLY_BOX) b2o_polygon_set_box(sh, d->p[0], d->p[1]);
                else if (d->shapeKind == B2G_POLY_BOX_OFFSET)
                    b2o_polygon_set_box_offset(sh, d->p[0], d->p[1], V2(d->center[0], d->center[1]), d->angle);
                else {
                    v2 pts[B2O_MAX_POLY];
                    int n = d->vertexCount < B2O_MAX_POLY ? d->vertexCount : B2O_MAX_POLY;
                    for (int k = 0; k < n; ++k) pts[k] = V2(V[2 * (d->firstVertex + k)], V[2 * (d->firstVertex + k) + 1]);
                    b2o_polygon_set(sh, pts, n);
                }
                break;
            case B2G_SHAPE_CHAIN: {
                sh->type = B2G_SHAPE_CHAIN;
                sh->radius = B2O_POLYGON_RADIUS;
                sh->chain = w->chainPool + chainUsed;
                for (int k = 0; k < d->vertexCount; ++k)
                    sh->chain[k] = V2(V[2 * (d->firstVertex + k)], V[2 * (d->firstVertex + k) + 1]);
                if (d->shapeKind == B2G_CHAIN_LOOP) { /* b2ChainShape::CreateLoop — b2ChainShape.cpp:41-61 */
                    sh->chain[d->vertexCount] = sh->chain[0];
                    sh->chainCount = d->vertexCount + 1;
                    sh->prev = sh->chain[sh->chainCount - 2];
                    sh->next = sh->chain[1];
                    sh->hasPrev = sh->hasNext = 1;
                } else {
                    sh->chainCount = d->vertexCount;
                    sh->hasPrev = d->hasVertex0; sh->hasNext = d->hasVertex3;
                    sh->prev = V2(d->vertex0[0], d->vertex0[1]);
                    sh->next = V2(d->vertex3[0], d->vertex3[1]);
                }
                chainUsed += sh->chainCount;
                childCount = sh->chainCount - 1;
            } break;
        }
        body_t* b = &w->bodies[f->body];
        f->firstProxy = w->np;
        f->proxyCount = 0;
        if (b->flags & BF_ACTIVE) {
            f->proxyCount = childCount;
            for (int k = 0; k < childCount; ++k) {
                proxy_t* p = &w->proxies[w->np];
                aabb t;
                b2o_shape_aabb(sh, b->xf, k, &t);
                v2 r = V2(B2O_AABB_EXTENSION, B2O_AABB_EXTENSION);
                p->fat.lo = v2_sub(t.lo, r);
                p->fat.hi = v2_add(t.hi, r);
                p->fixture = i;
                p->child = k;
                p->id = w->np == 0 ? 0 : 2 * w->np - 1; /* tree node id, SURVEY.md Appendix B.3 */
                buffer_move(w, w->np);
                w->np++;
            }
        }
        if (f->density > 0.0f) reset_mass_data(w, f->body);
        w->newFixture = 1;
    }
    w->islandBodies = (int*)calloc(4 * (w->nb + 1), sizeof(int));
    return w;
}

void b2o_world_destroy(void* h) {
    world_t* w = (world_t*)h;
    free(w->bodies); free(w->fixtures); free(w->proxies); free(w->contacts); free(w->chainPool);
    free(w->moveBuffer); free(w->events); free(w->islandBodies); free(w->islandContacts);
    free(w);
}

/* b2ContactFilter::ShouldCollide — Dynamics/b2WorldCallbacks.cpp:26-38 */
static int filter_should_collide(const fixture_t* a, const fixture_t* b) {
    if (a->group == b->group && a->group != 0) return a->group > 0;
    return (a->mask & b->cat) != 0 && (a->cat & b->mask) != 0;
}

/* b2Body::ShouldCollide — Dynamics/b2Body.cpp:395-416 (no joints in scope) */
static int body_should_collide(const body_t* a, const body_t* b) {
    if (a->type != B2G_DYNAMIC_BODY && b->type != B2G_DYNAMIC_BODY) return 0;
    return 1;
}

/* b2Contact::s_registers — Dynamics/Contacts/b2Contact.cpp:42-76.  returns 0 none, 1 primary, 2 swapped */
static int contact_register(int t1, int t2) {
    if (t1 == B2G_SHAPE_CIRCLE && t2 == B2G_SHAPE_CIRCLE) return 1;
    if (t1 == B2G_SHAPE_POLYGON && t2 == B2G_SHAPE_CIRCLE) return 1;
    if (t1 == B2G_SHAPE_CIRCLE && t2 == B2G_SHAPE_POLYGON) return 2;
    if (t1 == B2G_SHAPE_POLYGON && t2 == B2G_SHAPE_POLYGON) return 1;
    if (t1 == B2G_SHAPE_EDGE && t2 == B2G_SHAPE_CIRCLE) return 1;
    if (t1 == B2G_SHAPE_CIRCLE && t2 == B2G_SHAPE_EDGE) return 2;
    if (t1 == B2G_SHAPE_EDGE && t2 == B2G_SHAPE_POLYGON) return 1;
    if (t1 == B2G_SHAPE_POLYGON && t2 == B2G_SHAPE_EDGE) return 2;
    if (t1 == B2G_SHAPE_CHAIN && t2 == B2G_SHAPE_CIRCLE) return 1;
    if (t1 == B2G_SHAPE_CIRCLE && t2 == B2G_SHAPE_CHAIN) return 2;
    if (t1 == B2G_SHAPE_CHAIN && t2 == B2G_SHAPE_POLYGON) return 1;
    if (t1 == B2G_SHAPE_POLYGON && t2 == B2G_SHAPE_CHAIN) return 2;
    return 0;
}

/* b2ContactManager::AddPair — Dynamics/b2ContactManager.cpp:179-298 */
static void add_pair(world_t* w, int proxyA, int proxyB) {
    int fixtureA = w->proxies[proxyA].fixture, fixtureB = w->proxies[proxyB].fixture;
    int indexA = w->proxies[proxyA].child, indexB = w->proxies[proxyB].child;
    int bodyA = w->fixtures[fixtureA].body, bodyB = w->fixtures[fixtureB].body;
    if (bodyA == bodyB) return;
    for (int i = 0; i < w->nc; ++i) { /* does a contact already exist? */
        const contact_t* c = &w->contacts[i];
        if (c->fixtureA == fixtureA && c->fixtureB == fixtureB && c->childA == indexA && c->childB == indexB) return;
        if (c->fixtureA == fixtureB && c->fixtureB == fixtureA && c->childA == indexB && c->childB == indexA) return;
    }
    if (!body_should_collide(&w->bodies[bodyB], &w->bodies[bodyA])) return;
    if (!filter_should_collide(&w->fixtures[fixtureA], &w->fixtures[fixtureB])) return;
    /* b2Contact::Create — b2Contact.cpp:78-109 */
    int reg = contact_register(w->fixtures[fixtureA].shape.type, w->fixtures[fixtureB].shape.type);
    if (reg == 0) return;
    if (reg == 2) {
        int t;
        t = fixtureA; fixtureA = fixtureB; fixtureB = t;
        t = indexA; indexA = indexB; indexB = t;
        t = proxyA; proxyA = proxyB; proxyB = t;
        t = bodyA; bodyA = bodyB; bodyB = t;
    }
    if (w->nc == w->capc) {
        w->capc = w->capc ? 2 * w->capc : 256;
        w->contacts = (contact_t*)realloc(w->contacts, sizeof(contact_t) * w->capc);
    }
    contact_t* c = &w->contacts[w->nc++]; /* append = head-insert into all three lists */
    memset(c, 0, sizeof(*c));
    /* b2Contact::b2Contact — b2Contact.cpp:135-166; mixing b2Contact.h:50-63 */
    c->flags = CF_ENABLED;
    c->fixtureA = fixtureA; c->fixtureB = fixtureB;
    c->childA = indexA; c->childB = indexB;
    c->proxyA = proxyA; c->proxyB = proxyB;
    c->m.pointCount = 0;
    c->friction = sqrtf(w->fixtures[fixtureA].friction * w->fixtures[fixtureB].friction);
    c->restitution = w->fixtures[fixtureA].restitution > w->fixtures[fixtureB].restitution
                         ? w->fixtures[fixtureA].restitution
                         : w->fixtures[fixtureB].restitution;
    c->tangentSpeed = 0.0f;
    if (!w->fixtures[fixtureA].sensor && !w->fixtures[fixtureB].sensor) {
        set_awake(&w->bodies[bodyA], 1);
        set_awake(&w->bodies[bodyB], 1);
    }
}

typedef struct { int a, b; } pair_t; /* proxy indices ordered by proxy id */

static world_t* g_sort_world;
static int pair_less(const void* x, const void* y) { /* b2PairLessThan — Collision/b2BroadPhase.h:141-154 */
    const pair_t* p = (const pair_t*)x;
    const pair_t* q = (const pair_t*)y;
    int pa = g_sort_world->proxies[p->a].id, pb = g_sort_world->proxies[p->b].id;
    int qa = g_sort_world->proxies[q->a].id, qb = g_sort_world->proxies[q->b].id;
    if (pa != qa) return pa < qa ? -1 : 1;
    if (pb != qb) return pb < qb ? -1 : 1;
    return 0;
}

/* b2BroadPhase::UpdatePairs — Collision/b2BroadPhase.h:188-236 (+ QueryCallback b2BroadPhase.cpp:98-121) */
static void find_new_contacts(world_t* w) {
    int cap = 1024, n = 0;
    pair_t* pairs = (pair_t*)malloc(sizeof(pair_t) * cap);
    for (int i = 0; i < w->moveCount; ++i) {
        int q = w->moveBuffer[i];
        aabb fat = w->proxies[q].fat;
        for (int p = 0; p < w->np; ++p) {
            if (p == q) continue;
            if (!aabb_overlap(w->proxies[p].fat, fat)) continue;
            if (n == cap) { cap *= 2; pairs = (pair_t*)realloc(pairs, sizeof(pair_t) * cap); }
            if (w->proxies[p].id < w->proxies[q].id) { pairs[n].a = p; pairs[n].b = q; }
            else { pairs[n].a = q; pairs[n].b = p; }
            ++n;
        }
    }
    w->moveCount = 0;
    g_sort_world = w;
    qsort(pairs, n, sizeof(pair_t), pair_less);
    int i = 0;
    while (i < n) {
        pair_t primary = pairs[i];
        add_pair(w, primary.a, primary.b);
        ++i;
        while (i < n && pairs[i].a == primary.a && pairs[i].b == primary.b) ++i;
    }
    free(pairs);
}

/* b2Contact::Evaluate of the 7 subclasses — Dynamics/Contacts/b2*Contact.cpp */
static void evaluate(world_t* w, contact_t* c, xform xfA, xform xfB) {
    const b2o_shape* sA = &w->fixtures[c->fixtureA].shape;
    const b2o_shape* sB = &w->fixtures[c->fixtureB].shape;
    b2o_shape edge;
    if (sA->type == B2G_SHAPE_CHAIN) { b2o_chain_child_edge(sA, c->childA, &edge); sA = &edge; }
    if (sA->type == B2G_SHAPE_POLYGON && sB->type == B2G_SHAPE_POLYGON) b2o_collide_polygons_m(&c->m, sA, xfA, sB, xfB);
    else if (sA->type == B2G_SHAPE_CIRCLE && sB->type == B2G_SHAPE_CIRCLE) b2o_collide_circles_m(&c->m, sA, xfA, sB, xfB);
    else if (sA->type == B2G_SHAPE_POLYGON && sB->type == B2G_SHAPE_CIRCLE) b2o_collide_polygon_circle_m(&c->m, sA, xfA, sB, xfB);
    else if (sA->type == B2G_SHAPE_EDGE && sB->type == B2G_SHAPE_CIRCLE) b2o_collide_edge_circle_m(&c->m, sA, xfA, sB, xfB);
    else if (sA->type == B2G_SHAPE_EDGE && sB->type == B2G_SHAPE_POLYGON) b2o_collide_edge_polygon_m(&c->m, sA, xfA, sB, xfB);
}

/* b2ContactManager::Destroy — Dynamics/b2ContactManager.cpp:38-102 + b2Contact::Destroy b2Contact.cpp:111-133 */
static void destroy_contact(world_t* w, int ci) {
    contact_t* c = &w->contacts[ci];
    if (c->flags & CF_TOUCHING) push_event(w, 1, c, c->m.pointCount, 0, 0);
    if (c->m.pointCount > 0 && !w->fixtures[c->fixtureA].sensor && !w->fixtures[c->fixtureB].sensor) {
        set_awake(&w->bodies[w->fixtures[c->fixtureA].body], 1);
        set_awake(&w->bodies[w->fixtures[c->fixtureB].body], 1);
    }
    memmove(&w->contacts[ci], &w->contacts[ci + 1], sizeof(contact_t) * (w->nc - ci - 1));
    w->nc--;
}

/* b2Contact::Update — Dynamics/Contacts/b2Contact.cpp:168-251 (sensors are out of scope: DESIGN.md) */
static void update_contact(world_t* w, contact_t* c) {
    b2o_manifold oldManifold = c->m;
    c->flags |= CF_ENABLED;
    int touching = 0;
    int wasTouching = (c->flags & CF_TOUCHING) == CF_TOUCHING;
    body_t* bodyA = &w->bodies[w->fixtures[c->fixtureA].body];
    body_t* bodyB = &w->bodies[w->fixtures[c->fixtureB].body];
    evaluate(w, c, bodyA->xf, bodyB->xf);
    touching = c->m.pointCount > 0;
    for (int i = 0; i < c->m.pointCount; ++i) {
        b2o_mpoint* mp2 = c->m.points + i;
        mp2->normalImpulse = 0.0f;
        mp2->tangentImpulse = 0.0f;
        for (int j = 0; j < oldManifold.pointCount; ++j) {
            b2o_mpoint* mp1 = oldManifold.points + j;
            if (mp1->key == mp2->key) {
                mp2->normalImpulse = mp1->normalImpulse;
                mp2->tangentImpulse = mp1->tangentImpulse;
                break;
            }
        }
    }
    if (touching != wasTouching) { set_awake(bodyA, 1); set_awake(bodyB, 1); }
    if (touching) c->flags |= CF_TOUCHING; else c->flags &= ~CF_TOUCHING;
    if (!wasTouching && touching) push_event(w, 0, c, c->m.pointCount, 0, 0);
    if (wasTouching && !touching) push_event(w, 1, c, c->m.pointCount, 0, 0);
    if (touching) push_event(w, 2, c, c->m.pointCount, 0, 0);
}

/* b2ContactManager::Collide — Dynamics/b2ContactManager.cpp:107-172 */
static void collide(world_t* w) {
    for (int i = w->nc - 1; i >= 0; --i) {
        contact_t* c = &w->contacts[i];
        body_t* bodyA = &w->bodies[w->fixtures[c->fixtureA].body];
        body_t* bodyB = &w->bodies[w->fixtures[c->fixtureB].body];
        if (c->flags & CF_FILTER) {
            if (!body_should_collide(bodyB, bodyA)) { destroy_contact(w, i); continue; }
            if (!filter_should_collide(&w->fixtures[c->fixtureA], &w->fixtures[c->fixtureB])) { destroy_contact(w, i); continue; }
            c->flags &= ~CF_FILTER;
        }
        int activeA = (bodyA->flags & BF_AWAKE) && bodyA->type != B2G_STATIC_BODY;
        int activeB = (bodyB->flags & BF_AWAKE) && bodyB->type != B2G_STATIC_BODY;
        if (!activeA && !activeB) continue;
        if (!aabb_overlap(w->proxies[c->proxyA].fat, w->proxies[c->proxyB].fat)) { destroy_contact(w, i); continue; }
        update_contact(w, c);
    }
}

/* ---- island solver ------------------------------------------------------------------------- */

typedef struct { v2 rA, rB; float normalImpulse, tangentImpulse, normalMass, tangentMass, velocityBias; } vcp_t;
typedef struct { /* b2ContactVelocityConstraint — Dynamics/Contacts/b2ContactSolver.h:33-57 */
    vcp_t points[2];
    v2 normal;
    float nm_exx, nm_exy, nm_eyx, nm_eyy; /* normalMass (b2Mat22 ex, ey) */
    float K_exx, K_exy, K_eyx, K_eyy;
    int indexA, indexB;
    float invMassA, invMassB, invIA, invIB, friction, restitution, tangentSpeed;
    int pointCount, contactIndex;
} vc_t;
typedef struct { /* b2ContactPositionConstraint — b2ContactSolver.cpp:33-47 */
    v2 localPoints[2], localNormal, localPoint;
    int indexA, indexB;
    float invMassA, invMassB;
    v2 localCenterA, localCenterB;
    float invIA, invIB;
    int type;
    float radiusA, radiusB;
    int pointCount;
} pc_t;
typedef struct { v2 c; float a; } pos_t;
typedef struct { v2 v; float w; } vel_t;

/* b2WorldManifold::Initialize — Collision/b2Collision.cpp:25-87 */
static void world_manifold(const b2o_manifold* m, xform xfA, float radiusA, xform xfB, float radiusB, v2* normal,
                           v2 points[2]) {
    if (m->pointCount == 0) return;
    switch (m->type) {
        case B2O_MANIFOLD_CIRCLES: {
            *normal = V2(1.0f, 0.0f);
            v2 pointA = xf_mul(xfA, m->localPoint);
            v2 pointB = xf_mul(xfB, m->points[0].localPoint);
            if (v2_dist_sq(pointA, pointB) > B2O_EPSILON * B2O_EPSILON) {
                *normal = v2_sub(pointB, pointA);
                v2_normalize(normal);
            }
            v2 cA = v2_add(pointA, v2_scale(radiusA, *normal));
            v2 cB = v2_sub(pointB, v2_scale(radiusB, *normal));
            points[0] = v2_scale(0.5f, v2_add(cA, cB));
        } break;
        case B2O_MANIFOLD_FACE_A: {
            *normal = rot_mul(xfA.q, m->localNormal);
            v2 planePoint = xf_mul(xfA, m->localPoint);
            for (int i = 0; i < m->pointCount; ++i) {
                v2 clipPoint = xf_mul(xfB, m->points[i].localPoint);
                v2 cA = v2_add(clipPoint, v2_scale(radiusA - v2_dot(v2_sub(clipPoint, planePoint), *normal), *normal));
                v2 cB = v2_sub(clipPoint, v2_scale(radiusB, *normal));
                points[i] = v2_scale(0.5f, v2_add(cA, cB));
            }
        } break;
        case B2O_MANIFOLD_FACE_B: {
            *normal = rot_mul(xfB.q, m->localNormal);
            v2 planePoint = xf_mul(xfB, m->localPoint);
            for (int i = 0; i < m->pointCount; ++i) {
                v2 clipPoint = xf_mul(xfA, m->points[i].localPoint);
                v2 cB = v2_add(clipPoint, v2_scale(radiusB - v2_dot(v2_sub(clipPoint, planePoint), *normal), *normal));
                v2 cA = v2_sub(clipPoint, v2_scale(radiusA, *normal));
                points[i] = v2_scale(0.5f, v2_add(cA, cB));
            }
            *normal = v2_neg(*normal);
        } break;
    }
}

/* b2Island::Solve — Dynamics/b2Island.cpp:185-383 with b2ContactSolver — Contacts/b2ContactSolver.cpp */
static void solve_island(world_t* w, const int* ibodies, int nbody, const int* icontacts, int ncontact, float h,
                         float dtRatio, int velIters, int posIters) {
    pos_t* positions = (pos_t*)malloc(sizeof(pos_t) * (nbody + 1));
    vel_t* velocities = (vel_t*)malloc(sizeof(vel_t) * (nbody + 1));
    vc_t* vcs = (vc_t*)calloc(ncontact + 1, sizeof(vc_t));
    pc_t* pcs = (pc_t*)calloc(ncontact + 1, sizeof(pc_t));
    int warmStarting = (w->flags & B2G_WORLD_WARM_STARTING) != 0;

    /* integrate velocities — b2Island.cpp:190-225 */
    for (int i = 0; i < nbody; ++i) {
        body_t* b = &w->bodies[ibodies[i]];
        v2 c = b->c;
        float a = b->a;
        v2 v = b->v;
        float wv = b->w;
        b->c0 = b->c;
        b->a0 = b->a;
        if (b->type == B2G_DYNAMIC_BODY) {
            /* v += h * (gravityScale * gravity + invMass * force) */
            v = v2_add(v, v2_scale(h, v2_add(v2_scale(b->gravityScale, w->gravity), v2_scale(b->invMass, b->force))));
            wv += h * b->invI * b->torque;
            v = v2_scale(1.0f / (1.0f + h * b->linearDamping), v);
            wv *= 1.0f / (1.0f + h * b->angularDamping);
        }
        positions[i].c = c; positions[i].a = a;
        velocities[i].v = v; velocities[i].w = wv;
    }

    /* b2ContactSolver ctor — b2ContactSolver.cpp:49-136 */
    for (int i = 0; i < ncontact; ++i) {
        contact_t* contact = &w->contacts[icontacts[i]];
        fixture_t* fA = &w->fixtures[contact->fixtureA];
        fixture_t* fB = &w->fixtures[contact->fixtureB];
        body_t* bodyA = &w->bodies[fA->body];
        body_t* bodyB = &w->bodies[fB->body];
        b2o_manifold* manifold = &contact->m;
        int pointCount = manifold->pointCount;
        vc_t* vc = &vcs[i];
        vc->friction = contact->friction;
        vc->restitution = contact->restitution;
        vc->tangentSpeed = contact->tangentSpeed;
        vc->indexA = bodyA->islandIndex; vc->indexB = bodyB->islandIndex;
        vc->invMassA = bodyA->invMass; vc->invMassB = bodyB->invMass;
        vc->invIA = bodyA->invI; vc->invIB = bodyB->invI;
        vc->contactIndex = i;
        vc->pointCount = pointCount;
        pc_t* pc = &pcs[i];
        pc->indexA = bodyA->islandIndex; pc->indexB = bodyB->islandIndex;
        pc->invMassA = bodyA->invMass; pc->invMassB = bodyB->invMass;
        pc->localCenterA = bodyA->localCenter; pc->localCenterB = bodyB->localCenter;
        pc->invIA = bodyA->invI; pc->invIB = bodyB->invI;
        pc->localNormal = manifold->localNormal;
        pc->localPoint = manifold->localPoint;
        pc->pointCount = pointCount;
        pc->radiusA = fA->shape.radius; pc->radiusB = fB->shape.radius;
        pc->type = manifold->type;
        for (int j = 0; j < pointCount; ++j) {
            b2o_mpoint* cp = manifold->points + j;
            vcp_t* vcp = vc->points + j;
            if (warmStarting) {
                vcp->normalImpulse = dtRatio * cp->normalImpulse;
                vcp->tangentImpulse = dtRatio * cp->tangentImpulse;
            } else {
                vcp->normalImpulse = 0.0f;
                vcp->tangentImpulse = 0.0f;
            }
            pc->localPoints[j] = cp->localPoint;
        }
    }

    /* InitializeVelocityConstraints — b2ContactSolver.cpp:145-254 */
    for (int i = 0; i < ncontact; ++i) {
        vc_t* vc = &vcs[i];
        pc_t* pc = &pcs[i];
        float radiusA = pc->radiusA, radiusB = pc->radiusB;
        b2o_manifold* manifold = &w->contacts[icontacts[vc->contactIndex]].m;
        int indexA = vc->indexA, indexB = vc->indexB;
        float mA = vc->invMassA, mB = vc->invMassB, iA = vc->invIA, iB = vc->invIB;
        v2 localCenterA = pc->localCenterA, localCenterB = pc->localCenterB;
        v2 cA = positions[indexA].c; float aA = positions[indexA].a;
        v2 vA = velocities[indexA].v; float wA = velocities[indexA].w;
        v2 cB = positions[indexB].c; float aB = positions[indexB].a;
        v2 vB = velocities[indexB].v; float wB = velocities[indexB].w;
        xform xfA, xfB;
        xfA.q = rot_set(aA);
        xfB.q = rot_set(aB);
        xfA.p = v2_sub(cA, rot_mul(xfA.q, localCenterA));
        xfB.p = v2_sub(cB, rot_mul(xfB.q, localCenterB));
        v2 wmPoints[2];
        v2 wmNormal = V2(0, 0);
        world_manifold(manifold, xfA, radiusA, xfB, radiusB, &wmNormal, wmPoints);
        vc->normal = wmNormal;
        int pointCount = vc->pointCount;
        for (int j = 0; j < pointCount; ++j) {
            vcp_t* vcp = vc->points + j;
            vcp->rA = v2_sub(wmPoints[j], cA);
            vcp->rB = v2_sub(wmPoints[j], cB);
            float rnA = v2_cross(vcp->rA, vc->normal);
            float rnB = v2_cross(vcp->rB, vc->normal);
            float kNormal = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
            vcp->normalMass = kNormal > 0.0f ? 1.0f / kNormal : 0.0f;
            v2 tangent = v2_cross_vs(vc->normal, 1.0f);
            float rtA = v2_cross(vcp->rA, tangent);
            float rtB = v2_cross(vcp->rB, tangent);
            float kTangent = mA + mB + iA * rtA * rtA + iB * rtB * rtB;
            vcp->tangentMass = kTangent > 0.0f ? 1.0f / kTangent : 0.0f;
            vcp->velocityBias = 0.0f;
            /* vRel = dot(normal, vB + cross(wB, rB) - vA - cross(wA, rA)) */
            v2 dv = v2_sub(v2_sub(v2_add(vB, v2_cross_sv(wB, vcp->rB)), vA), v2_cross_sv(wA, vcp->rA));
            float vRel = v2_dot(vc->normal, dv);
            if (vRel < -B2O_VELOCITY_THRESHOLD) vcp->velocityBias = -vc->restitution * vRel;
        }
        if (vc->pointCount == 2) {
            vcp_t* vcp1 = vc->points + 0;
            vcp_t* vcp2 = vc->points + 1;
            float rn1A = v2_cross(vcp1->rA, vc->normal);
            float rn1B = v2_cross(vcp1->rB, vc->normal);
            float rn2A = v2_cross(vcp2->rA, vc->normal);
            float rn2B = v2_cross(vcp2->rB, vc->normal);
            float k11 = mA + mB + iA * rn1A * rn1A + iB * rn1B * rn1B;
            float k22 = mA + mB + iA * rn2A * rn2A + iB * rn2B * rn2B;
            float k12 = mA + mB + iA * rn1A * rn2A + iB * rn1B * rn2B;
            const float k_maxConditionNumber = 1000.0f;
            if (k11 * k11 < k_maxConditionNumber * (k11 * k22 - k12 * k12)) {
                vc->K_exx = k11; vc->K_exy = k12;
                vc->K_eyx = k12; vc->K_eyy = k22;
                /* b2Mat22::GetInverse — b2Math.h:211-225 */
                float a = vc->K_exx, b = vc->K_eyx, c = vc->K_exy, d = vc->K_eyy;
                float det = a * d - b * c;
                if (det != 0.0f) det = 1.0f / det;
                vc->nm_exx = det * d; vc->nm_eyx = -det * b;
                vc->nm_exy = -det * c; vc->nm_eyy = det * a;
            } else {
                vc->pointCount = 1;
            }
        }
    }

    /* WarmStart — b2ContactSolver.cpp:256-294 */
    if (warmStarting) {
        for (int i = 0; i < ncontact; ++i) {
            vc_t* vc = &vcs[i];
            int indexA = vc->indexA, indexB = vc->indexB;
            float mA = vc->invMassA, iA = vc->invIA, mB = vc->invMassB, iB = vc->invIB;
            v2 vA = velocities[indexA].v; float wA = velocities[indexA].w;
            v2 vB = velocities[indexB].v; float wB = velocities[indexB].w;
            v2 normal = vc->normal;
            v2 tangent = v2_cross_vs(normal, 1.0f);
            for (int j = 0; j < vc->pointCount; ++j) {
                vcp_t* vcp = vc->points + j;
                v2 P = v2_add(v2_scale(vcp->normalImpulse, normal), v2_scale(vcp->tangentImpulse, tangent));
                wA -= iA * v2_cross(vcp->rA, P);
                vA = v2_sub(vA, v2_scale(mA, P));
                wB += iB * v2_cross(vcp->rB, P);
                vB = v2_add(vB, v2_scale(mB, P));
            }
            velocities[indexA].v = vA; velocities[indexA].w = wA;
            velocities[indexB].v = vB; velocities[indexB].w = wB;
        }
    }

    /* SolveVelocityConstraints — b2ContactSolver.cpp:296-613 */
    for (int it = 0; it < velIters; ++it) {
        for (int i = 0; i < ncontact; ++i) {
            vc_t* vc = &vcs[i];
            int indexA = vc->indexA, indexB = vc->indexB;
            float mA = vc->invMassA, iA = vc->invIA, mB = vc->invMassB, iB = vc->invIB;
            int pointCount = vc->pointCount;
            v2 vA = velocities[indexA].v; float wA = velocities[indexA].w;
            v2 vB = velocities[indexB].v; float wB = velocities[indexB].w;
            v2 normal = vc->normal;
            v2 tangent = v2_cross_vs(normal, 1.0f);
            float friction = vc->friction;
            for (int j = 0; j < pointCount; ++j) {
                vcp_t* vcp = vc->points + j;
                v2 dv = v2_sub(v2_sub(v2_add(vB, v2_cross_sv(wB, vcp->rB)), vA), v2_cross_sv(wA, vcp->rA));
                float vt = v2_dot(dv, tangent) - vc->tangentSpeed;
                float lambda = vcp->tangentMass * (-vt);
                float maxFriction = friction * vcp->normalImpulse;
                float newImpulse = f_clamp(vcp->tangentImpulse + lambda, -maxFriction, maxFriction);
                lambda = newImpulse - vcp->tangentImpulse;
                vcp->tangentImpulse = newImpulse;
                v2 P = v2_scale(lambda, tangent);
                vA = v2_sub(vA, v2_scale(mA, P));
                wA -= iA * v2_cross(vcp->rA, P);
                vB = v2_add(vB, v2_scale(mB, P));
                wB += iB * v2_cross(vcp->rB, P);
            }
            if (pointCount == 1) {
                for (int j = 0; j < pointCount; ++j) {
                    vcp_t* vcp = vc->points + j;
                    v2 dv = v2_sub(v2_sub(v2_add(vB, v2_cross_sv(wB, vcp->rB)), vA), v2_cross_sv(wA, vcp->rA));
                    float vn = v2_dot(dv, normal);
                    float lambda = -vcp->normalMass * (vn - vcp->velocityBias);
                    float newImpulse = f_max(vcp->normalImpulse + lambda, 0.0f);
                    lambda = newImpulse - vcp->normalImpulse;
                    vcp->normalImpulse = newImpulse;
                    v2 P = v2_scale(lambda, normal);
                    vA = v2_sub(vA, v2_scale(mA, P));
                    wA -= iA * v2_cross(vcp->rA, P);
                    vB = v2_add(vB, v2_scale(mB, P));
                    wB += iB * v2_cross(vcp->rB, P);
                }
            } else {
                vcp_t* cp1 = vc->points + 0;
                vcp_t* cp2 = vc->points + 1;
                v2 a = V2(cp1->normalImpulse, cp2->normalImpulse);
                v2 dv1 = v2_sub(v2_sub(v2_add(vB, v2_cross_sv(wB, cp1->rB)), vA), v2_cross_sv(wA, cp1->rA));
                v2 dv2 = v2_sub(v2_sub(v2_add(vB, v2_cross_sv(wB, cp2->rB)), vA), v2_cross_sv(wA, cp2->rA));
                float vn1 = v2_dot(dv1, normal);
                float vn2 = v2_dot(dv2, normal);
                v2 b;
                b.x = vn1 - cp1->velocityBias;
                b.y = vn2 - cp2->velocityBias;
                /* b -= b2Mul(K, a) */
                b.x -= vc->K_exx * a.x + vc->K_eyx * a.y;
                b.y -= vc->K_exy * a.x + vc->K_eyy * a.y;
                v2 x, d;
                int solved = 0;
                for (;;) {
                    /* case 1 */
                    x.x = -(vc->nm_exx * b.x + vc->nm_eyx * b.y);
                    x.y = -(vc->nm_exy * b.x + vc->nm_eyy * b.y);
                    if (x.x >= 0.0f && x.y >= 0.0f) { solved = 1; break; }
                    /* case 2 */
                    x.x = -cp1->normalMass * b.x;
                    x.y = 0.0f;
                    vn1 = 0.0f;
                    vn2 = vc->K_exy * x.x + b.y;
                    if (x.x >= 0.0f && vn2 >= 0.0f) { solved = 1; break; }
                    /* case 3 */
                    x.x = 0.0f;
                    x.y = -cp2->normalMass * b.y;
                    vn1 = vc->K_eyx * x.y + b.x;
                    vn2 = 0.0f;
                    if (x.y >= 0.0f && vn1 >= 0.0f) { solved = 1; break; }
                    /* case 4 */
                    x.x = 0.0f;
                    x.y = 0.0f;
                    vn1 = b.x;
                    vn2 = b.y;
                    if (vn1 >= 0.0f && vn2 >= 0.0f) { solved = 1; break; }
                    break;
                }
                if (solved) {
                    d = v2_sub(x, a);
                    v2 P1 = v2_scale(d.x, normal);
                    v2 P2 = v2_scale(d.y, normal);
                    vA = v2_sub(vA, v2_scale(mA, v2_add(P1, P2)));
                    wA -= iA * (v2_cross(cp1->rA, P1) + v2_cross(cp2->rA, P2));
                    vB = v2_add(vB, v2_scale(mB, v2_add(P1, P2)));
                    wB += iB * (v2_cross(cp1->rB, P1) + v2_cross(cp2->rB, P2));
                    cp1->normalImpulse = x.x;
                    cp2->normalImpulse = x.y;
                }
            }
            velocities[indexA].v = vA; velocities[indexA].w = wA;
            velocities[indexB].v = vB; velocities[indexB].w = wB;
        }
    }

    /* StoreImpulses — b2ContactSolver.cpp:615-628 */
    for (int i = 0; i < ncontact; ++i) {
        vc_t* vc = &vcs[i];
        b2o_manifold* manifold = &w->contacts[icontacts[vc->contactIndex]].m;
        for (int j = 0; j < vc->pointCount; ++j) {
            manifold->points[j].normalImpulse = vc->points[j].normalImpulse;
            manifold->points[j].tangentImpulse = vc->points[j].tangentImpulse;
        }
    }

    /* integrate positions — b2Island.cpp:276-306 */
    for (int i = 0; i < nbody; ++i) {
        v2 c = positions[i].c; float a = positions[i].a;
        v2 v = velocities[i].v; float wv = velocities[i].w;
        v2 translation = v2_scale(h, v);
        if (v2_dot(translation, translation) > B2O_MAX_TRANSLATION_SQ) {
            float ratio = B2O_MAX_TRANSLATION / v2_len(translation);
            v = v2_scale(ratio, v);
        }
        float rotation = h * wv;
        if (rotation * rotation > B2O_MAX_ROTATION_SQ) {
            float ratio = B2O_MAX_ROTATION / f_abs(rotation);
            wv *= ratio;
        }
        c = v2_add(c, v2_scale(h, v));
        a += h * wv;
        positions[i].c = c; positions[i].a = a;
        velocities[i].v = v; velocities[i].w = wv;
    }

    /* SolvePositionConstraints — b2Island.cpp:308-328, b2ContactSolver.cpp:630-760 */
    int positionSolved = 0;
    for (int it = 0; it < posIters; ++it) {
        float minSeparation = 0.0f;
        for (int i = 0; i < ncontact; ++i) {
            pc_t* pc = &pcs[i];
            int indexA = pc->indexA, indexB = pc->indexB;
            v2 localCenterA = pc->localCenterA, localCenterB = pc->localCenterB;
            float mA = pc->invMassA, iA = pc->invIA, mB = pc->invMassB, iB = pc->invIB;
            int pointCount = pc->pointCount;
            v2 cA = positions[indexA].c; float aA = positions[indexA].a;
            v2 cB = positions[indexB].c; float aB = positions[indexB].a;
            for (int j = 0; j < pointCount; ++j) {
                xform xfA, xfB;
                xfA.q = rot_set(aA);
                xfB.q = rot_set(aB);
                xfA.p = v2_sub(cA, rot_mul(xfA.q, localCenterA));
                xfB.p = v2_sub(cB, rot_mul(xfB.q, localCenterB));
                /* b2PositionSolverManifold::Initialize — b2ContactSolver.cpp:630-683 */
                v2 normal, point;
                float separation;
                switch (pc->type) {
                    case B2O_MANIFOLD_CIRCLES: {
                        v2 pointA = xf_mul(xfA, pc->localPoint);
                        v2 pointB = xf_mul(xfB, pc->localPoints[0]);
                        normal = v2_sub(pointB, pointA);
                        v2_normalize(&normal);
                        point = v2_scale(0.5f, v2_add(pointA, pointB));
                        separation = v2_dot(v2_sub(pointB, pointA), normal) - pc->radiusA - pc->radiusB;
                    } break;
                    case B2O_MANIFOLD_FACE_A: {
                        normal = rot_mul(xfA.q, pc->localNormal);
                        v2 planePoint = xf_mul(xfA, pc->localPoint);
                        v2 clipPoint = xf_mul(xfB, pc->localPoints[j]);
                        separation = v2_dot(v2_sub(clipPoint, planePoint), normal) - pc->radiusA - pc->radiusB;
                        point = clipPoint;
                    } break;
                    default: {
                        normal = rot_mul(xfB.q, pc->localNormal);
                        v2 planePoint = xf_mul(xfB, pc->localPoint);
                        v2 clipPoint = xf_mul(xfA, pc->localPoints[j]);
                        separation = v2_dot(v2_sub(clipPoint, planePoint), normal) - pc->radiusA - pc->radiusB;
                        point = clipPoint;
                        normal = v2_neg(normal);
                    } break;
                }
                v2 rA = v2_sub(point, cA);
                v2 rB = v2_sub(point, cB);
                minSeparation = f_min(minSeparation, separation);
                float C = f_clamp(B2O_BAUMGARTE * (separation + B2O_LINEAR_SLOP), -B2O_MAX_LINEAR_CORRECTION, 0.0f);
                float rnA = v2_cross(rA, normal);
                float rnB = v2_cross(rB, normal);
                float K = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
                float impulse = K > 0.0f ? -C / K : 0.0f;
                v2 P = v2_scale(impulse, normal);
                cA = v2_sub(cA, v2_scale(mA, P));
                aA -= iA * v2_cross(rA, P);
                cB = v2_add(cB, v2_scale(mB, P));
                aB += iB * v2_cross(rB, P);
            }
            positions[indexA].c = cA; positions[indexA].a = aA;
            positions[indexB].c = cB; positions[indexB].a = aB;
        }
        int contactsOkay = minSeparation >= -3.0f * B2O_LINEAR_SLOP;
        if (contactsOkay) { positionSolved = 1; break; }
    }

    /* copy back + SynchronizeTransform — b2Island.cpp:330-339, b2Body.h:835-839 */
    for (int i = 0; i < nbody; ++i) {
        body_t* b = &w->bodies[ibodies[i]];
        b->c = positions[i].c;
        b->a = positions[i].a;
        b->v = velocities[i].v;
        b->w = velocities[i].w;
        b->xf.q = rot_set(b->a);
        b->xf.p = v2_sub(b->c, rot_mul(b->xf.q, b->localCenter));
    }

    /* Report → PostSolve — b2Island.cpp:517-540 */
    for (int i = 0; i < ncontact; ++i) {
        vc_t* vc = &vcs[i];
        float ni[2] = {0, 0}, ti[2] = {0, 0};
        for (int j = 0; j < vc->pointCount; ++j) { ni[j] = vc->points[j].normalImpulse; ti[j] = vc->points[j].tangentImpulse; }
        push_event(w, 3, &w->contacts[icontacts[i]], vc->pointCount, ni, ti);
    }

    /* sleep — b2Island.cpp:345-382 */
    if (w->flags & B2G_WORLD_ALLOW_SLEEP) {
        float minSleepTime = B2O_MAXFLOAT;
        const float linTolSqr = B2O_LINEAR_SLEEP_TOL * B2O_LINEAR_SLEEP_TOL;
        const float angTolSqr = B2O_ANGULAR_SLEEP_TOL * B2O_ANGULAR_SLEEP_TOL;
        for (int i = 0; i < nbody; ++i) {
            body_t* b = &w->bodies[ibodies[i]];
            if (b->type == B2G_STATIC_BODY) continue;
            if ((b->flags & BF_AUTOSLEEP) == 0 || b->w * b->w > angTolSqr || v2_dot(b->v, b->v) > linTolSqr) {
                b->sleepTime = 0.0f;
                minSleepTime = 0.0f;
            } else {
                b->sleepTime += h;
                minSleepTime = f_min(minSleepTime, b->sleepTime);
            }
        }
        if (minSleepTime >= B2O_TIME_TO_SLEEP && positionSolved) {
            for (int i = 0; i < nbody; ++i) set_awake(&w->bodies[ibodies[i]], 0);
        }
    }
    free(positions); free(velocities); free(vcs); free(pcs);
}

/* b2World::Solve — Dynamics/b2World.cpp:388-576 */
static void solve(world_t* w, float dt, float dtRatio, int velIters, int posIters) {
    for (int i = 0; i < w->nb; ++i) w->bodies[i].flags &= ~BF_ISLAND;
    for (int i = 0; i < w->nc; ++i) w->contacts[i].flags &= ~CF_ISLAND;
    int* stack = (int*)malloc(sizeof(int) * (w->nb + 1));
    int* ibodies = (int*)malloc(sizeof(int) * (w->nb + 1));
    int* icontacts = (int*)malloc(sizeof(int) * (w->nc + 1));
    free(w->islandContacts);
    w->islandContacts = (int*)calloc(w->nc + 1, sizeof(int));
    w->islandBodyCount = w->islandContactCount = w->islandCount = 0;

    for (int seed = w->nb - 1; seed >= 0; --seed) { /* body list: newest first */
        body_t* sb = &w->bodies[seed];
        if (sb->flags & BF_ISLAND) continue;
        if ((sb->flags & BF_AWAKE) == 0 || (sb->flags & BF_ACTIVE) == 0) continue;
        if (sb->type == B2G_STATIC_BODY) continue;
        int nbody = 0, ncontact = 0, stackCount = 0;
        stack[stackCount++] = seed;
        sb->flags |= BF_ISLAND;
        while (stackCount > 0) {
            int bi = stack[--stackCount];
            body_t* b = &w->bodies[bi];
            b->islandIndex = nbody;
            ibodies[nbody++] = bi;
            set_awake(b, 1);
            if (b->type == B2G_STATIC_BODY) continue;
            for (int ci = w->nc - 1; ci >= 0; --ci) { /* this body's edge list = world list filtered */
                contact_t* c = &w->contacts[ci];
                int bodyA = w->fixtures[c->fixtureA].body, bodyB = w->fixtures[c->fixtureB].body;
                if (bodyA != bi && bodyB != bi) continue;
                if (c->flags & CF_ISLAND) continue;
                if ((c->flags & CF_ENABLED) == 0 || (c->flags & CF_TOUCHING) == 0) continue;
                if (w->fixtures[c->fixtureA].sensor || w->fixtures[c->fixtureB].sensor) continue;
                icontacts[ncontact++] = ci;
                c->flags |= CF_ISLAND;
                int other = bodyA == bi ? bodyB : bodyA;
                if (w->bodies[other].flags & BF_ISLAND) continue;
                stack[stackCount++] = other;
                w->bodies[other].flags |= BF_ISLAND;
            }
        }
        if (w->islandBodyCount + nbody <= 4 * (w->nb + 1)) {
            memcpy(w->islandBodies + w->islandBodyCount, ibodies, sizeof(int) * nbody);
            w->islandBodyCount += nbody;
        }
        memcpy(w->islandContacts + w->islandContactCount, icontacts, sizeof(int) * ncontact);
        w->islandContactCount += ncontact;
        w->islandCount++;
        solve_island(w, ibodies, nbody, icontacts, ncontact, dt, dtRatio, velIters, posIters);
        for (int i = 0; i < nbody; ++i) {
            body_t* b = &w->bodies[ibodies[i]];
            if (b->type == B2G_STATIC_BODY) b->flags &= ~BF_ISLAND;
        }
    }
    free(stack); free(ibodies); free(icontacts);

    /* SynchronizeFixtures — b2World.cpp:551-569, b2Body.cpp:442-453, b2Fixture.cpp:154-177,
     * b2DynamicTree::MoveProxy — Collision/b2DynamicTree.cpp:133-177 */
    for (int bi = w->nb - 1; bi >= 0; --bi) {
        body_t* b = &w->bodies[bi];
        if ((b->flags & BF_ISLAND) == 0) continue;
        if (b->type == B2G_STATIC_BODY) continue;
        xform xf1;
        xf1.q = rot_set(b->a0);
        xf1.p = v2_sub(b->c0, rot_mul(xf1.q, b->localCenter));
        for (int fi = w->nf - 1; fi >= 0; --fi) {
            fixture_t* f = &w->fixtures[fi];
            if (f->body != bi) continue;
            for (int k = 0; k < f->proxyCount; ++k) {
                proxy_t* p = &w->proxies[f->firstProxy + k];
                aabb aabb1, aabb2, box;
                b2o_shape_aabb(&f->shape, xf1, p->child, &aabb1);
                b2o_shape_aabb(&f->shape, b->xf, p->child, &aabb2);
                box.lo = v2_min(aabb1.lo, aabb2.lo);
                box.hi = v2_max(aabb1.hi, aabb2.hi);
                v2 displacement = v2_sub(b->xf.p, xf1.p);
                if (aabb_contains(p->fat, box)) continue;
                v2 r = V2(B2O_AABB_EXTENSION, B2O_AABB_EXTENSION);
                box.lo = v2_sub(box.lo, r);
                box.hi = v2_add(box.hi, r);
                v2 d = v2_scale(B2O_AABB_MULTIPLIER, displacement);
                if (d.x < 0.0f) box.lo.x += d.x; else box.hi.x += d.x;
                if (d.y < 0.0f) box.lo.y += d.y; else box.hi.y += d.y;
                p->fat = box;
                buffer_move(w, f->firstProxy + k);
            }
        }
    }
    find_new_contacts(w);
}

/* b2World::Step — Dynamics/b2World.cpp:899-963 (continuous physics / TOI out of scope: DESIGN.md) */
void b2o_world_step(void* h, float dt, int velIters, int posIters) {
    world_t* w = (world_t*)h;
    w->eventCount = 0;
    if (w->newFixture) {
        find_new_contacts(w);
        w->newFixture = 0;
    }
    float inv_dt = dt > 0.0f ? 1.0f / dt : 0.0f;
    float dtRatio = w->inv_dt0 * dt;
    collide(w);
    if (dt > 0.0f) solve(w, dt, dtRatio, velIters, posIters);
    if (dt > 0.0f) w->inv_dt0 = inv_dt;
    if (w->flags & B2G_WORLD_CLEAR_FORCES) { /* ClearForces — b2World.cpp:965-972 */
        for (int i = 0; i < w->nb; ++i) {
            w->bodies[i].force = V2(0.0f, 0.0f);
            w->bodies[i].torque = 0.0f;
        }
    }
}

/* Between-step body edits (ops = B2GPU_CMD_* of include/b2gpu.h) — Box2D/Dynamics/b2Body.h:738-833 (Apply*),
 * :408-467 (SetLinearVelocity / SetAngularVelocity), :634-653 (SetAwake), b2Body.cpp:418-440 (SetTransform). */
void b2o_world_body_command(void* h, int body, int op, int wake, float vx, float vy, float px, float py) {
    world_t* w = (world_t*)h;
    if (body < 0 || body >= w->nb) return;
    body_t* b = &w->bodies[body];
    v2 v = V2(vx, vy), p = V2(px, py);
    switch (op) {
        case 0: case 1: case 2:
            if (b->type != B2G_DYNAMIC_BODY) return;
            if (wake && (b->flags & BF_AWAKE) == 0) set_awake(b, 1);
            if ((b->flags & BF_AWAKE) == 0) return;
            if (op == 2) { b->torque += vx; return; }
            b->force = v2_add(b->force, v);
            if (op == 0) b->torque += v2_cross(v2_sub(p, b->c), v);
            return;
        case 3: case 4:
            if (b->type != B2G_DYNAMIC_BODY) return;
            if (wake && (b->flags & BF_AWAKE) == 0) set_awake(b, 1);
            if ((b->flags & BF_AWAKE) == 0) return;
            if (op == 4) { b->w += b->invI * vx; return; }
            b->v = v2_add(b->v, v2_scale(b->invMass, v));
            b->w += b->invI * v2_cross(v2_sub(p, b->c), v);
            return;
        case 5:
            if (b->type == B2G_STATIC_BODY) return;
            if (v2_dot(v, v) > 0.0f) set_awake(b, 1);
            b->v = v;
            return;
        case 6:
            if (b->type == B2G_STATIC_BODY) return;
            if (vx * vx > 0.0f) set_awake(b, 1);
            b->w = vx;
            return;
        case 8:
            set_awake(b, wake);
            return;
        case 7: {
            b->xf.q = rot_set(px);
            b->xf.p = v;
            b->c = xf_mul(b->xf, b->localCenter);
            b->a = px;
            b->c0 = b->c;
            b->a0 = px;
            for (int fi = w->nf - 1; fi >= 0; --fi) {
                fixture_t* f = &w->fixtures[fi];
                if (f->body != body) continue;
                for (int k = 0; k < f->proxyCount; ++k) {
                    proxy_t* pr = &w->proxies[f->firstProxy + k];
                    aabb a1, box;
                    b2o_shape_aabb(&f->shape, b->xf, pr->child, &a1);
                    box.lo = v2_min(a1.lo, a1.lo);
                    box.hi = v2_max(a1.hi, a1.hi);
                    v2 displacement = v2_sub(b->xf.p, b->xf.p);
                    if (aabb_contains(pr->fat, box)) continue;
                    v2 r = V2(B2O_AABB_EXTENSION, B2O_AABB_EXTENSION);
                    box.lo = v2_sub(box.lo, r);
                    box.hi = v2_add(box.hi, r);
                    v2 d = v2_scale(B2O_AABB_MULTIPLIER, displacement);
                    if (d.x < 0.0f) box.lo.x += d.x; else box.hi.x += d.x;
                    if (d.y < 0.0f) box.lo.y += d.y; else box.hi.y += d.y;
                    pr->fat = box;
                    buffer_move(w, f->firstProxy + k);
                }
            }
            return;
        }
    }
}

/* ---- dumps: same formats as oracle/ref_harness.cpp ------------------------------------------ */
void b2o_world_record_events(void* h, int on) { ((world_t*)h)->recordEvents = on; }
int b2o_world_body_count(void* h) { return ((world_t*)h)->nb; }
int b2o_world_contact_count(void* h) { return ((world_t*)h)->nc; }
int b2o_world_proxy_count(void* h) { return ((world_t*)h)->np; }

void b2o_world_get_bodies(void* h, float* out) {
    world_t* w = (world_t*)h;
    for (int i = 0; i < w->nb; ++i) {
        body_t* b = &w->bodies[i];
        float* o = out + 16 * i;
        o[0] = b->xf.p.x; o[1] = b->xf.p.y; o[2] = b->xf.q.s; o[3] = b->xf.q.c;
        o[4] = b->c.x; o[5] = b->c.y; o[6] = b->a;
        o[7] = b->v.x; o[8] = b->v.y; o[9] = b->w;
        o[10] = (b->flags & BF_AWAKE) ? 1.0f : 0.0f;
        o[11] = b->mass;
        o[12] = b->I + b->mass * v2_dot(b->localCenter, b->localCenter); /* b2Body::GetInertia */
        o[13] = b->localCenter.x; o[14] = b->localCenter.y;
        o[15] = 0.0f;
    }
}

int b2o_world_get_contacts(void* h, int32_t* ints, float* floats, int cap) {
    world_t* w = (world_t*)h;
    int n = 0;
    for (int ci = w->nc - 1; ci >= 0; --ci, ++n) {
        if (n >= cap) continue;
        contact_t* c = &w->contacts[ci];
        int32_t* I = ints + 8 * n;
        float* F = floats + 16 * n;
        I[0] = c->fixtureA; I[1] = c->childA; I[2] = c->fixtureB; I[3] = c->childB;
        I[4] = ((c->flags & CF_TOUCHING) ? 1 : 0) | ((c->flags & CF_ENABLED) ? 2 : 0);
        I[5] = c->m.pointCount;
        I[6] = c->m.pointCount > 0 ? (int32_t)c->m.points[0].key : 0;
        I[7] = c->m.pointCount > 1 ? (int32_t)c->m.points[1].key : 0;
        memset(F, 0, 16 * sizeof(float));
        if (c->m.pointCount > 0) {
            F[0] = c->m.localNormal.x; F[1] = c->m.localNormal.y;
            F[2] = c->m.localPoint.x; F[3] = c->m.localPoint.y;
            F[4] = (float)c->m.type;
            for (int k = 0; k < c->m.pointCount; ++k) {
                F[5 + 4 * k] = c->m.points[k].localPoint.x;
                F[6 + 4 * k] = c->m.points[k].localPoint.y;
                F[7 + 4 * k] = c->m.points[k].normalImpulse;
                F[8 + 4 * k] = c->m.points[k].tangentImpulse;
            }
        }
        F[13] = c->friction;
        F[14] = c->restitution;
    }
    return n;
}

int b2o_world_get_proxies(void* h, float* aabbs, int32_t* ids, int cap) {
    world_t* w = (world_t*)h;
    for (int i = 0; i < w->np && i < cap; ++i) {
        aabbs[4 * i] = w->proxies[i].fat.lo.x; aabbs[4 * i + 1] = w->proxies[i].fat.lo.y;
        aabbs[4 * i + 2] = w->proxies[i].fat.hi.x; aabbs[4 * i + 3] = w->proxies[i].fat.hi.y;
        ids[i] = w->proxies[i].id;
    }
    return w->np;
}

int b2o_world_get_events(void* h, int32_t* ints, float* floats, int cap) {
    world_t* w = (world_t*)h;
    for (int i = 0; i < w->eventCount && i < cap; ++i) {
        const event_t* e = &w->events[i];
        int32_t* I = ints + 6 * i;
        I[0] = e->type; I[1] = e->fixtureA; I[2] = e->childA; I[3] = e->fixtureB; I[4] = e->childB; I[5] = e->count;
        float* F = floats + 4 * i;
        F[0] = e->normalImpulses[0]; F[1] = e->normalImpulses[1];
        F[2] = e->tangentImpulses[0]; F[3] = e->tangentImpulses[1];
    }
    return w->eventCount;
}

/* island order of the last step: bodies then contacts ((fixtureA, fixtureB) per contact) */
int b2o_world_get_islands(void* h, int32_t* bodies, int bodyCap, int32_t* counts) {
    world_t* w = (world_t*)h;
    for (int i = 0; i < w->islandBodyCount && i < bodyCap; ++i) bodies[i] = w->islandBodies[i];
    counts[0] = w->islandCount; counts[1] = w->islandBodyCount; counts[2] = w->islandContactCount;
    return w->islandBodyCount;
}

/* ---- stateless known-answer entry points (same signatures as ref_harness.cpp) --------------- */
int b2o_polygon_build(const b2g_fixture_def* fd, const float* verts, float* out) {
    b2o_shape p;
    memset(&p, 0, sizeof(p));
    if (fd->shapeKind == B2G_POLY_BOX) b2o_polygon_set_box(&p, fd->p[0], fd->p[1]);
    else if (fd->shapeKind == B2G_POLY_BOX_OFFSET)
        b2o_polygon_set_box_offset(&p, fd->p[0], fd->p[1], V2(fd->center[0], fd->center[1]), fd->angle);
    else {
        v2 pts[B2O_MAX_POLY];
        int n = fd->vertexCount < B2O_MAX_POLY ? fd->vertexCount : B2O_MAX_POLY;
        for (int i = 0; i < n; ++i) pts[i] = V2(verts[2 * (fd->firstVertex + i)], verts[2 * (fd->firstVertex + i) + 1]);
        b2o_polygon_set(&p, pts, n);
    }
    for (int i = 0; i < 8; ++i) {
        out[2 * i] = i < p.count ? p.verts[i].x : 0.0f;
        out[2 * i + 1] = i < p.count ? p.verts[i].y : 0.0f;
        out[16 + 2 * i] = i < p.count ? p.normals[i].x : 0.0f;
        out[16 + 2 * i + 1] = i < p.count ? p.normals[i].y : 0.0f;
    }
    out[32] = p.centroid.x; out[33] = p.centroid.y;
    b2o_massdata md;
    b2o_shape_mass(&p, fd->density, &md);
    out[34] = md.mass; out[35] = md.center.x; out[36] = md.center.y; out[37] = md.I;
    return p.count;
}

static void fill_poly(b2o_shape* p, const float* poly) {
    memset(p, 0, sizeof(*p));
    p->type = B2G_SHAPE_POLYGON;
    p->radius = B2O_POLYGON_RADIUS;
    p->count = (int)poly[0];
    for (int i = 0; i < 8; ++i) {
        p->verts[i] = V2(poly[1 + 2 * i], poly[2 + 2 * i]);
        p->normals[i] = V2(poly[17 + 2 * i], poly[18 + 2 * i]);
    }
    p->centroid = V2(poly[33], poly[34]);
}
static void fill_circle(b2o_shape* c, const float* f) {
    memset(c, 0, sizeof(*c));
    c->type = B2G_SHAPE_CIRCLE;
    c->p = V2(f[0], f[1]);
    c->radius = f[2];
}
static void fill_edge(b2o_shape* e, const float* f) {
    memset(e, 0, sizeof(*e));
    e->type = B2G_SHAPE_EDGE;
    e->radius = B2O_POLYGON_RADIUS;
    e->v0 = V2(f[0], f[1]); e->v1 = V2(f[2], f[3]); e->v2_ = V2(f[4], f[5]); e->v3 = V2(f[6], f[7]);
    e->has0 = f[8] != 0.0f; e->has3 = f[9] != 0.0f;
}
static xform mk_xf(const float* x) { xform t; t.p = V2(x[0], x[1]); t.q.s = x[2]; t.q.c = x[3]; return t; }
static void dump_manifold(const b2o_manifold* m, int32_t* ints, float* floats) {
    ints[0] = m->pointCount;
    ints[1] = m->type;
    ints[2] = m->pointCount > 0 ? (int32_t)m->points[0].key : 0;
    ints[3] = m->pointCount > 1 ? (int32_t)m->points[1].key : 0;
    memset(floats, 0, 8 * sizeof(float));
    if (m->pointCount > 0) {
        floats[0] = m->localNormal.x; floats[1] = m->localNormal.y;
        floats[2] = m->localPoint.x; floats[3] = m->localPoint.y;
        for (int k = 0; k < m->pointCount; ++k) {
            floats[4 + 2 * k] = m->points[k].localPoint.x;
            floats[5 + 2 * k] = m->points[k].localPoint.y;
        }
    }
}
void b2o_collide_polygons(const float* a, const float* xa, const float* b, const float* xb, int32_t* ints, float* fl) {
    b2o_shape A, B; b2o_manifold m; memset(&m, 0, sizeof(m));
    fill_poly(&A, a); fill_poly(&B, b);
    b2o_collide_polygons_m(&m, &A, mk_xf(xa), &B, mk_xf(xb));
    dump_manifold(&m, ints, fl);
}
void b2o_collide_circles(const float* a, const float* xa, const float* b, const float* xb, int32_t* ints, float* fl) {
    b2o_shape A, B; b2o_manifold m; memset(&m, 0, sizeof(m));
    fill_circle(&A, a); fill_circle(&B, b);
    b2o_collide_circles_m(&m, &A, mk_xf(xa), &B, mk_xf(xb));
    dump_manifold(&m, ints, fl);
}
void b2o_collide_polygon_circle(const float* a, const float* xa, const float* b, const float* xb, int32_t* ints, float* fl) {
    b2o_shape A, B; b2o_manifold m; memset(&m, 0, sizeof(m));
    fill_poly(&A, a); fill_circle(&B, b);
    b2o_collide_polygon_circle_m(&m, &A, mk_xf(xa), &B, mk_xf(xb));
    dump_manifold(&m, ints, fl);
}
void b2o_collide_edge_circle(const float* a, const float* xa, const float* b, const float* xb, int32_t* ints, float* fl) {
    b2o_shape A, B; b2o_manifold m; memset(&m, 0, sizeof(m));
    fill_edge(&A, a); fill_circle(&B, b);
    b2o_collide_edge_circle_m(&m, &A, mk_xf(xa), &B, mk_xf(xb));
    dump_manifold(&m, ints, fl);
}
void b2o_collide_edge_polygon(const float* a, const float* xa, const float* b, const float* xb, int32_t* ints, float* fl) {
    b2o_shape A, B; b2o_manifold m; memset(&m, 0, sizeof(m));
    fill_edge(&A, a); fill_poly(&B, b);
    b2o_collide_edge_polygon_m(&m, &A, mk_xf(xa), &B, mk_xf(xb));
    dump_manifold(&m, ints, fl);
}
