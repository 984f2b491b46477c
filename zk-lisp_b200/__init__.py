"""zk-lisp_b200 — B200-native (sm_100a CUDA) implementation of the Winterfell proving hot path that
yoozzeek/zk-lisp drives once per execution segment (zk-lisp-proof-winterfell/src/prove.rs:425-517).

The directory name carries a hyphen (task layout); import it through `import_zkl()` in
`__graft_entry__.py`, which registers it as the module `zk_lisp_b200`.
"""
from . import agg, agg_child, air, air_eval, capi, field, params, planner, preflight, recursion, replay, segments, sharded, sharding, step, verify  # noqa: F401
from .prover import Context, ZkProver, hasher_suite  # noqa: F401
