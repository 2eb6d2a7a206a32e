"""Lowering of a host-side world (rooms / entities built by `world.MiniWorldEnv`) into the
flat records of include/mwb.h.  Pure data movement: every number is taken from the same
numpy arrays the reference would have handed to OpenGL (wall_verts, wall_norms, texcoords,
miniworld.py:286-434) or to its collision code (wall_segs, :325)."""
import numpy as np

from .assets import Texture
from .engine import ENTITY_DTYPE, MAX_EDGES, PROTO_DTYPE, QUAD_DTYPE, ROOM_DTYPE, SEG_DTYPE, SURF_CEIL, \
    SURF_FLOOR, SURF_WALL
from .entity import KIND_AGENT, KIND_BOX, KIND_MESH, Agent, Box, MeshEnt


def room_cdf(room_probs):
    """The cdf numpy's Generator.choice(n, p=p) searches: p.cumsum() / p.cumsum()[-1]."""
    cdf = np.asarray(room_probs, dtype=np.float64).cumsum()
    cdf /= cdf[-1]
    return cdf


def pack_geometry(env):
    """rooms -> (mwb_room[], mwb_quad[], mwb_seg[]), in the reference's draw / list order."""
    rooms = np.zeros(len(env.rooms), ROOM_DTYPE)
    cdf = room_cdf(env.room_probs)
    quads = []
    for ri, r in enumerate(env.rooms):
        rec = rooms[ri]
        rec["min_x"], rec["max_x"], rec["min_z"], rec["max_z"] = r.min_x, r.max_x, r.min_z, r.max_z
        rec["cdf"] = cdf[ri]
        n = r.num_walls
        if n > MAX_EDGES:
            raise ValueError("room outline has %d vertices (max %d)" % (n, MAX_EDGES))
        rec["num_edges"] = n
        rec["edge_px"][:n], rec["edge_pz"][:n] = r.outline[:, 0], r.outline[:, 2]
        rec["edge_nx"][:n], rec["edge_nz"][:n] = r.edge_norms[:, 0], r.edge_norms[:, 2]
        for k, (name, tex) in enumerate(((r.wall_tex_name, r.wall_tex), (r.floor_tex_name, r.floor_tex),
                                         (r.ceil_tex_name, r.ceil_tex))):
            fam = Texture.family(name)
            rec["tex_first"][k], rec["tex_count"][k], rec["tex_id"][k] = fam[0].tex_id, len(fam), tex.tex_id

        def polygon(verts, surf, normal):
            # GL_POLYGON -> triangle fan; packed as quads (0, k, k+1, k+2) / a final triangle
            nv = len(verts)
            k = 1
            while k + 1 < nv:
                take = 3 if k + 2 < nv else 2
                idx = [0] + list(range(k, k + take))
                q = np.zeros((), QUAD_DTYPE)
                q["num_verts"] = len(idx)
                for j, vi in enumerate(idx):
                    q["pos"][j] = verts[vi]
                    q["uvm"][j] = (verts[vi][0], verts[vi][2])
                q["nrm"], q["room"], q["surf"] = normal, ri, surf
                quads.append(q)
                k += take - 1

        polygon(r.floor_verts, SURF_FLOOR, (0, 1, 0))
        if not r.no_ceiling:
            polygon(r.ceil_verts, SURF_CEIL, (0, -1, 0))
        for wi in range(len(r.wall_verts) // 4):
            q = np.zeros((), QUAD_DTYPE)
            q["num_verts"] = 4
            q["pos"] = r.wall_verts[4 * wi:4 * wi + 4]
            q["uvm"] = r.wall_uvm[4 * wi:4 * wi + 4]
            q["nrm"], q["room"], q["surf"] = r.wall_norms[4 * wi], ri, SURF_WALL
            quads.append(q)
    quads = np.array(quads, QUAD_DTYPE) if quads else np.zeros(0, QUAD_DTYPE)
    ws = np.asarray(env.wall_segs, dtype=np.float64).reshape(-1, 2, 3)
    segs = np.zeros(len(ws), SEG_DTYPE)
    segs["ax"], segs["az"], segs["bx"], segs["bz"] = ws[:, 0, 0], ws[:, 0, 2], ws[:, 1, 0], ws[:, 1, 2]
    return rooms, quads, segs


def proto_record(ent):
    """Episode-constant description of an entity (mwb_proto)."""
    p = np.zeros((), PROTO_DTYPE)
    p["mesh_id"] = -1
    p["radius"] = float(ent.radius)
    p["radius_is_f32"] = int(isinstance(ent.radius, np.float32))
    p["height"] = float(ent.height)
    p["is_static"] = int(bool(ent.is_static))
    if isinstance(ent, Agent):
        p["kind"] = KIND_AGENT
    elif isinstance(ent, Box):
        p["kind"] = KIND_BOX
        p["size"] = np.asarray(ent.size, float)
        from .entity import COLORS
        p["color"] = COLORS[ent.color]
    elif isinstance(ent, MeshEnt) or getattr(ent, "mesh", None) is not None:   # OBJ meshes and quad frames
        p["kind"] = KIND_MESH
        p["mesh_id"] = ent.mesh.mesh_id
        p["scale"] = np.float32(ent.scale)
        p["deg_form"] = int(isinstance(ent, MeshEnt))     # glRotatef angle: dir * 180 / pi vs dir * (180 / pi)
    else:
        raise TypeError("entity type %s is not supported by the CUDA engine yet" % type(ent).__name__)
    return p


def pack_world(env):
    """Everything mwb_set_world needs for one host-generated env."""
    rooms, quads, segs = pack_geometry(env)
    slot_entities = list(env.entities)
    carried = env.agent.carrying
    if carried is not None and carried not in slot_entities:
        slot_entities.append(carried)
    protos = np.zeros(len(slot_entities), PROTO_DTYPE)
    ents = np.zeros(len(slot_entities), ENTITY_DTYPE)
    agent_slot = -1
    for k, ent in enumerate(slot_entities):
        protos[k] = proto_record(ent)
        ents[k]["proto"] = k
        ents[k]["pos"] = np.asarray(ent.pos, float)
        ents[k]["dir"] = float(ent.dir)
        ents[k]["color"] = np.asarray(getattr(ent, "color_vec", (1.0, 1.0, 1.0)), float)
        if ent is env.agent:
            agent_slot = k
    a = env.agent
    return dict(rooms=rooms, quads=quads, segs=segs, protos=protos, ents=ents, slot_entities=slot_entities,
                agent_slot=agent_slot, carrying=slot_entities.index(carried) if carried is not None else -1,
                step_count=env.step_count, num_picked_up=getattr(env, "num_picked_up", 0),
                cam=(a.cam_height, a.cam_fwd_disp, a.cam_pitch, a.cam_fov_y),
                sky_color=env.sky_color, light_pos=env.light_pos, light_color=env.light_color,
                light_ambient=env.light_ambient)
