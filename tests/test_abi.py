"""CPU-side checks of the boundary: the shared library builds/loads, exports every symbol include/fql_b200.h declares,
and its host-only entry points (layout, sizes, validation, error reporting) behave.  No compute calls (no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def L():
    import __graft_entry__ as g
    g.build()
    from fql_b200 import _lib
    return _lib


def test_every_declared_symbol_is_exported(L):
    hdr = open(os.path.join(ROOT, 'include', 'fql_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = set(re.findall(r'\b(fql_[a-z_0-9]+)\s*\(', hdr))
    assert len(declared) >= 20
    lib = C.CDLL(L.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in include/fql_b200.h but not exported'
    assert declared == set(L.EXPORTS), declared ^ set(L.EXPORTS)


def test_struct_sizes_match_header(L):
    assert C.sizeof(L.FqlDims) == 16 * 4
    assert C.sizeof(L.FqlHparams) == 12 * 4
    assert C.sizeof(L.FqlBatch) == 10 * 8
    assert C.sizeof(L.FqlState) == 6 * 8
    assert C.sizeof(L.FqlLeaf) == 6 * 4 + 8


@pytest.mark.parametrize('F,A,trainable,target', [(28, 5, 3236364, 1619970), (29, 8, 3247634, 1624066), (69, 21, 3369516, 1678338)])
def test_layout_matches_reference_param_counts(L, F, A, trainable, target):
    """SURVEY 8a parameter counts of BASELINE configs 1-3 and the Flax leaf names/shapes."""
    d = L.make_dims(256, F, A)
    leaves, arena = L.layout(d)
    n = lambda lf: lf['ens'] * lf['rows'] * lf['cols']
    assert sum(n(l) for l in leaves if l['net'] != 'target_critic') == trainable
    assert sum(n(l) for l in leaves if l['net'] == 'target_critic') == target
    assert len(leaves) == 56                                    # 10 + 10 + 18 + 18
    first = {l['net']: l for l in reversed(leaves) if l['module'] == 'Dense_0' and l['is_kernel']}
    assert (first['actor_bc_flow']['rows'], first['actor_onestep_flow']['rows'], first['critic']['rows']) == (F + A + 1, F + A, F + A)
    assert first['critic']['ens'] == 2 and first['actor_bc_flow']['ens'] == 1
    offs = [l['offset'] for l in leaves]
    assert offs == sorted(offs) and all(o % 1024 == 0 for o in offs) and arena % 1024 == 0
    for a, b in zip(leaves, leaves[1:]):
        assert a['offset'] + n(a) <= b['offset']
    assert not any(l['module'].startswith('LayerNorm') for l in leaves if l['net'].startswith('actor'))
    d2 = L.make_dims(256, F, A, actor_layer_norm=True)
    assert len(L.layout(d2)[0]) == 72


def test_validation_errors_are_reported_not_thrown(L):
    lib = L.lib()
    bad = L.make_dims(0, 4, 2)
    assert lib.fql_arena_floats(C.byref(bad)) == -1
    assert b'batch' in lib.fql_last_error()
    # the peer-memory data-parallel interface: sizes and argument validation (no GPU needed)
    d = L.make_dims(128, 29, 8, global_batch=256, normalize_q_loss=True)
    arena = lib.fql_arena_floats(C.byref(d))
    nb = lib.fql_dp_symmetric_bytes(C.byref(d), 2)
    assert arena * 4 < nb < arena * 4 + (1 << 16) and nb % 256 == 0     # the gradient arena + gather / flag pads
    assert lib.fql_dp_symmetric_bytes(C.byref(d), 9) == 0 and b'world' in lib.fql_last_error()
    assert lib.fql_dp_attach(None, C.byref(d), None) != 0 and b'context' in lib.fql_last_error()
    with pytest.raises(L.FqlError):
        L.check(lib.fql_layout(C.byref(bad), None, 0, None), 'fql_layout')
    assert [lib.fql_info_name(i).decode() for i in range(13)] == list(__import__('fql_b200').INFO_KEYS)
    assert lib.fql_info_name(13) is None


def test_workspace_scales_with_batch_and_seeds(L):
    lib = L.lib()
    w = lambda **k: lib.fql_workspace_bytes(C.byref(L.make_dims(k.pop('batch', 256), 29, 8, **k)))
    assert w() < w(batch=512) < w(batch=1024)
    assert abs(w(num_seeds=4) / w(batch=1024) - 1) < 0.05      # seeds are one more batch-like axis


def test_get_config_matches_reference_defaults():
    from fql_b200 import get_config
    c = get_config()
    ref = dict(agent_name='fql', lr=3e-4, batch_size=256, actor_hidden_dims=(512,) * 4, value_hidden_dims=(512,) * 4, layer_norm=True,
               actor_layer_norm=False, discount=0.99, tau=0.005, q_agg='mean', alpha=300.0, flow_steps=10, normalize_q_loss=False,
               encoder=None)                                     # agents/fql.py:249-270
    for k, v in ref.items():
        assert tuple(c[k]) == v if isinstance(v, tuple) else c[k] == v, k
    assert c['actor_start_steps'] is None                       # main.py:198 reads it every iteration
    c.alpha = 10.0
    assert c['alpha'] == 10.0


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under fql_b200/ may reference it."""
    for dp, _, fs in os.walk(os.path.join(ROOT, 'fql_b200')):
        for f in fs:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', src, flags=re.M), f
