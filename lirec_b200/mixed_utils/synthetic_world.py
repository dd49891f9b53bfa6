"""A small synthetic MovieGraphs-like ANNOTATION world (the real annotations + 80 GB feature dump are
not available offline).

`build_world` produces what the reference's `MixedFeaturesDataset.__init__` obtains from its loaders
(`load_interaction_names`, `load_merged_interactions`, `load_set`, `load_annotated_inter`,
`load_iou2_clips`; mixed_utils/classification_dataloader.py:30-54, 108) plus the pooled feature
vectors the reference caches per interaction and per (interaction, character)
(mixed_utils/mixed_features.py:37-112):

  interactions   list of duck-typed `AnnotatedInter` objects (utils/util_functions.py:79-239) — only
                 the attributes the dataloader reads: id, inter_node, video_descr, time_node,
                 id2names, name2id, triplets, bi, ftracks, relships, get_relship_by_id
  rels           {movie: {(name1, name2): Relationship}} built with the Relationship API
                 (utils/util_functions.py:53-75) the way add_relationships does (:179-228)
  clip_vec       {inter_id: [1, text_dim + visual_dim]}
  track_vec      {(inter_id, name): [1, track_dim]} (zeros for a character without a face track)

The classes of the interaction and relationship objects are parameters, so the golden generator
(tests/golden/make_dataloader_golden.py) builds the SAME world out of the reference's own classes
and feeds it to the reference's unmodified dataset code, while the product builds it out of its own.
All feature values are small multiples of 1/8, exactly representable in bf16.
"""
from collections import defaultdict

import numpy as np

REL_NAMES = ["friend", "parent", "child", "colleague", "lover", "enemy", "sibling", "boss", "worker",
             "stranger", "customer", "teacher", "student", "neighbour", "ex-lover"]
REL_OPP = {"friend": "friend", "parent": "child", "child": "parent", "colleague": "colleague", "lover": "lover",
           "enemy": "enemy", "sibling": "sibling", "boss": "worker", "worker": "boss", "stranger": "stranger",
           "customer": "customer", "teacher": "student", "student": "teacher", "neighbour": "neighbour",
           "ex-lover": "ex-lover"}


class Inter:
    """Duck-typed AnnotatedInter (reference: utils/util_functions.py:79-97, 234-239)."""

    def get_relship_by_id(self, triplet_id):
        if triplet_id in self.relships:
            return np.random.choice(self.relships[triplet_id])
        return "None"


class World:
    pass


def _vec(rng, n, nonneg):
    v = rng.randint(-24, 25, size=(1, n)).astype(np.float64) / 8.0
    return np.abs(v) if nonneg else v


def build_world(seed=0, inter_cls=Inter, rel_cls=None, n_movies=3, n_scenes=7, n_chars=6, n_inter_names=40,
                n_merged=12, text_dim=768, visual_dim=2048, track_dim=2048, max_chars_per_inter=5,
                p_rel_node=0.45, p_empty_track=0.15):
    if rel_cls is None:
        from lirec_b200.utils.util_functions import Relationship as rel_cls
    rng = np.random.RandomState(seed)
    w = World()
    w.text_dim, w.visual_dim, w.track_dim = text_dim, visual_dim, track_dim

    # ---- interaction vocabulary: name -> (global idx, class type t/v/m, index inside the type) ----------
    names = ["inter%02d" % i for i in range(n_inter_names)]
    per_type = defaultdict(list)
    w.inter2idx = {}
    for g, nm in enumerate(names):
        t = int(rng.randint(3))
        w.inter2idx[nm] = (g, t, len(per_type[t]))
        per_type[t].append(nm)
    w.interaction_names = {"all": names, "t": per_type[0], "v": per_type[1], "m": per_type[2]}
    merged = ["mgd%02d" % i for i in range(n_merged)]
    w.inter2mgd = {nm: merged[int(rng.randint(n_merged))] for nm in names}
    w.mgd2idx = {m: i for i, m in enumerate(merged)}
    w.rels_opp = dict(REL_OPP)

    movies = ["tt%07d" % (100 + m) for m in range(n_movies)]
    w.split = {"train": movies[:max(1, n_movies - 2)], "val": movies[-2:-1] or movies[:1], "test": movies[-1:]}
    w.interactions, w.rels, w.clip_vec, w.track_vec = [], {}, {}, {}
    w.iou2_clips = defaultdict(lambda: defaultdict(list))
    w.by_movie = defaultdict(list)
    used_rels = set()

    for movie in movies:
        chars = ["%s char%d" % (movie[-2:], c) for c in range(n_chars)]
        dict_rel = {}
        for scene in range(1, n_scenes + 1):
            for _ in range(int(rng.randint(1, 4))):
                it = inter_cls.__new__(inter_cls)
                it.id = len(w.interactions)
                nm = names[int(rng.randint(len(names)))]
                it.inter_node = {"name": nm, "type": "interaction"}
                it.video_descr = {"movie": movie, "scene": [scene], "fname": ["%s.scene-%03d" % (movie, scene)]}
                it.time_node = {"start": float(rng.randint(10)), "end": float(10 + rng.randint(10)), "type": "time"}
                k = int(rng.randint(1, max_chars_per_inter + 1))
                present = [chars[i] for i in sorted(rng.choice(n_chars, size=k, replace=False))]
                it.id2names = {100 + i: p for i, p in enumerate(present)}
                it.name2id = {p: i for i, p in it.id2names.items()}
                it.paired_names, it.ftracks_names = {}, {}
                it.bi = False
                it.triplets, it.relships = {}, {}
                # triplets: directed pairs (a -> b), both directions when bidirectional, or one person
                if k >= 2 and rng.rand() < 0.85:
                    n_pairs = 1 + int(rng.rand() < 0.25 and k >= 3)
                    cnt = 0
                    for _p in range(n_pairs):
                        a, b = [present[i] for i in rng.choice(k, size=2, replace=False)]
                        it.bi = bool(rng.rand() < 0.35)
                        it.triplets[cnt] = {0: a, 1: b}
                        cnt += 1
                        if it.bi:
                            it.triplets[cnt] = {0: b, 1: a}
                            cnt += 1
                else:
                    a = present[int(rng.randint(k))]
                    it.bi = bool(rng.rand() < 0.3)
                    it.triplets[0] = {int(rng.randint(2)): a}
                    if rng.rand() < 0.2:
                        it.triplets[1] = {int(rng.randint(2)): a}
                it.tripl_counter = len(it.triplets)
                it.ftracks = defaultdict(list)
                for p in present:
                    empty = rng.rand() < p_empty_track
                    it.ftracks[p] = [] if empty else [{"frame": int(f), "timestamp": float(f) / 24.0}
                                                      for f in range(int(rng.randint(1, 5)))]
                    tv = np.zeros((1, track_dim)) if empty else _vec(rng, track_dim, True)
                    w.track_vec[(it.id, p)] = tv
                w.clip_vec[it.id] = np.hstack((_vec(rng, text_dim, False), _vec(rng, visual_dim, True)))

                # relationship nodes of the clip graph (add_relationships, util_functions.py:179-228)
                for tid, tr in it.triplets.items():
                    if len(tr) != 2 or rng.rand() >= p_rel_node:
                        continue
                    n1, n2 = tr[0], tr[1]
                    rel = REL_NAMES[int(rng.randint(len(REL_NAMES)))]
                    used_rels.update((rel, REL_OPP[rel]))
                    if (n1, n2) in dict_rel:
                        dict_rel[(n1, n2)].append_scene(rel, scene)
                        dict_rel[(n2, n1)].append_scene(REL_OPP[rel], scene)
                    else:
                        dict_rel[(n1, n2)] = rel_cls(rel, scene)
                        dict_rel[(n2, n1)] = rel_cls(REL_OPP[rel], scene)
                for r in dict_rel.values():
                    if scene not in r.scenes:
                        r.append_scene(rels_name=None, scene_idx=scene)
                for tid, tr in it.triplets.items():
                    if len(tr) == 2 and (tr[0], tr[1]) in dict_rel:
                        it.relships[tid] = dict_rel[(tr[0], tr[1])]._scene2rel[scene]
                # interactions of the same clip that overlap in time (soft labels / multilab weights)
                soft = [names[int(i)] for i in rng.choice(len(names), size=int(rng.randint(0, 4)), replace=False)]
                w.iou2_clips[(movie, scene)][nm] = soft
                w.interactions.append(it)
                w.by_movie[movie].append(it.id)
        w.rels[movie] = dict_rel
    w.rels_list = sorted(used_rels) + ["None"]
    return w


def subset(world, mode, inter_cls=None):
    """What `load_annotated_inter(movie_idxs=load_set(mode))` returns for one split: the split's
    interactions re-numbered from 0 (ids index the list), its relationship dicts, the relationship
    vocabulary and the opposite-relationship map."""
    ids = [i for m in world.split[mode] for i in world.by_movie[m]]
    remap = {old: new for new, old in enumerate(ids)}
    inters = []
    clip_vec, track_vec = {}, {}
    for old in ids:
        src = world.interactions[old]
        it = src.__class__.__new__(src.__class__)
        it.__dict__.update(src.__dict__)
        it.relships = dict(src.relships)
        it.id = remap[old]
        inters.append(it)
        clip_vec[it.id] = world.clip_vec[old]
        for p in it.id2names.values():
            track_vec[(it.id, p)] = world.track_vec[(old, p)]
    rels = {m: world.rels[m] for m in world.split[mode]}
    return dict(interactions=inters, rels=rels, rels_list=list(world.rels_list), rels_opp=dict(world.rels_opp),
                clip_vec=clip_vec, track_vec=track_vec, movie_idxs=list(world.split[mode]))
