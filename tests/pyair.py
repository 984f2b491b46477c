"""The Python-integer AIR lives in the product now (zk-lisp_b200/air_eval.py, used by the host verifier); the tests keep the short name."""
from zk_lisp_b200.air_eval import *  # noqa: F401,F403
from zk_lisp_b200.air_eval import assertions, evaluate_transition, merkle_block, poseidon_block, ram_block, rom_block  # noqa: F401
