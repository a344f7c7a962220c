import torch
from torch import nn

from deepspeed_b200.runtime.zero.units import arena_segments, build_units, param_fragments, PARAM_ALIGN, SHARD_ALIGN
from deepspeed_b200.utils import groups


class Block(nn.Module):

    def __init__(self, d):
        super().__init__()
        self.a = nn.Linear(d, d)
        self.n = nn.LayerNorm(d)


class Net(nn.Module):

    def __init__(self, d=24, L=3):
        super().__init__()
        self.emb = nn.Embedding(50, d)
        self.layers = nn.ModuleList([Block(d) for _ in range(L)])
        self.head = nn.Linear(d, 7)


def test_units_follow_module_list_and_cover_every_param():
    m = Net()
    for world in (1, 2, 3, 8):
        units = build_units(m, world)
        assert [u.name for u in units] == ["emb", "layers.0", "layers.1", "layers.2", "head"]
        seen = set()
        for u in units:
            assert u.full_numel % (world * SHARD_ALIGN) == 0 and u.shard_numel * world == u.full_numel
            prev_end = 0
            for s in u.slots:
                assert s.offset % PARAM_ALIGN == 0 and s.offset >= prev_end
                prev_end = s.offset + s.numel
                seen.add(id(s.param))
        assert seen == {id(p) for p in m.parameters()}
        assert units[-1].arena_offset + units[-1].shard_numel == sum(u.shard_numel for u in units)


def test_fragments_tile_each_param_exactly_once():
    m = Net()
    world = 4
    units = build_units(m, world)
    for u in units:
        for s in u.slots:
            frags = param_fragments(u, s, world)
            assert sum(f[3] for f in frags) == s.numel
            pos = 0
            for (_, p0, _, ln) in frags:
                assert p0 == pos
                pos += ln


def test_segments_cover_arena_and_respect_groups():
    m = Net()
    decay = [p for n, p in m.named_parameters() if p.ndim > 1]
    p2g = {id(p): (0 if any(p is q for q in decay) else 1) for p in m.parameters()}
    world = 2
    units = build_units(m, world, p2g)
    for rank in range(world):
        segs = arena_segments(units, rank)
        total = sum(u.shard_numel for u in units)
        assert segs[0].start == 0 and segs[-1].end == total
        for a, b in zip(segs, segs[1:]):
            assert a.end == b.start and a.group != b.group
        # every parameter element that lives on this rank is inside a segment of its own group
        for u in units:
            for s in u.slots:
                for (r, p0, a0, ln) in param_fragments(u, s, world):
                    if r != rank:
                        continue
                    hit = [g for g in segs if g.start <= a0 and a0 + ln <= g.end]
                    assert hit and hit[0].group == s.group


def test_tied_parameters_are_placed_once():
    m = Net()
    m.head2 = nn.Linear(24, 50, bias=False)
    m.head2.weight = m.emb.weight
    units = build_units(m, 2)
    n = sum(1 for u in units for s in u.slots if s.param is m.emb.weight)
    assert n == 1


def test_rank_layout_grid():
    lay = groups.rank_layout(16, tp=2, pp=2, sp=2)
    assert lay["tp"][0] == [0, 1] and lay["sp"][0] == [0, 2] and lay["dp"][0] == [0, 4]
    assert lay["pp"][0] == [0, 8]
    assert sorted(sum(lay["sdp"], [])) == list(range(16))
    ep, edp = groups.expert_layout([0, 1, 2, 3, 4, 5, 6, 7], 4)
    assert ep == [[0, 1, 2, 3], [4, 5, 6, 7]] and edp == [[0, 4], [1, 5], [2, 6], [3, 7]]
    ep, edp = groups.expert_layout([0, 1, 2, 3, 4, 5, 6, 7], 4, data_before_expert=True)
    assert ep == [[0, 2, 4, 6], [1, 3, 5, 7]] and edp == [[0, 1], [2, 3], [4, 5], [6, 7]]
