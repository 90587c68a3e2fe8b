"""CPU checks of the drop-in layer's CONTRACT: (1) every reference attribute / method name the stand-ins carry (tests/standins.CONTRACT)
occurs in the unmodified reference file it is cited from -- run where /root/reference exists; (2) every attribute the mixins read from
`self` is either defined by the mixin, listed in the contract, or private to the mixin (`_pulse*`) -- so the stand-ins cannot silently
drift from what the mixins need."""
import ast
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("PULSE_REFERENCE_ROOT", "/root/reference")


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
def test_contract_names_exist_in_the_reference_sources():
    from tests.standins import CONTRACT
    missing = []
    for side in CONTRACT.values():
        for rel, names in side.items():
            src = open(os.path.join(REF, rel)).read()
            for n in names:
                if not re.search(r"\b" + re.escape(n) + r"\b", src):
                    missing.append((rel, n))
    assert not missing, missing


def _self_attrs(path, class_name):
    tree = ast.parse(open(path).read())
    cls = next(n for n in ast.walk(tree) if isinstance(n, ast.ClassDef) and n.name == class_name)
    reads, defined = set(), {n.name for n in cls.body if isinstance(n, ast.FunctionDef)}
    for node in ast.walk(cls):
        if isinstance(node, ast.Attribute) and isinstance(node.value, ast.Name) and node.value.id == "self":
            reads.add(node.attr)
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Name) and node.func.id == "getattr" and len(node.args) >= 2 \
                and isinstance(node.args[0], ast.Name) and node.args[0].id == "self" and isinstance(node.args[1], ast.Constant):
            reads.add(node.args[1].value)
    return reads, defined


@pytest.mark.parametrize("path,cls,side", [("pulse_b200/humanoid_im.py", "HumanoidImB200Mixin", "task"),
                                           ("pulse_b200/agent_mixins.py", "AMPAgentB200Mixin", "agent")])
def test_mixins_only_touch_contract_names(path, cls, side):
    from tests.standins import CONTRACT
    reads, defined = _self_attrs(os.path.join(ROOT, path), cls)
    allowed = set(n for names in CONTRACT[side].values() for n in names) | defined
    extra = {"device", "num_envs", "dt", "vec_env", "model", "optimizer", "obs_shape", "actions_num", "add_obs_noise", "_pulse"}   # generic BaseTask / A2CBase fields
    unknown = sorted(a for a in reads if a not in allowed and a not in extra and not a.startswith("_pulse"))
    assert not unknown, f"{cls} reads attributes that are not in tests/standins.CONTRACT: {unknown}"
