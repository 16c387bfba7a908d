"""CPU: the drop-in scripts declare every command-line flag of the reference's scripts with the same option strings,
type and default (tests/golden/cli_contract.json is extracted from the reference's argparse sections by
oracle/gen_golden.py::cli_contract)."""
import importlib.util
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "links-3d-human-pose-estimation_b200")
with open(os.path.join(ROOT, "tests", "golden", "cli_contract.json")) as f:
    CONTRACT = json.load(f)


@pytest.mark.parametrize("script", sorted(CONTRACT))
def test_script_flags_match_reference(script):
    spec = importlib.util.spec_from_file_location("dropin_" + script[:-3], os.path.join(PKG, script))
    mod = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, [script]
    try:
        spec.loader.exec_module(mod)          # parsers are module-level; nothing runs outside __main__
    finally:
        sys.argv = argv
    actions = {o: a for a in mod.parser._actions for o in a.option_strings}
    for flag in CONTRACT[script]:
        for o in flag["options"]:
            assert o in actions, (script, o)
        a = actions[flag["options"][-1]]
        assert set(flag["options"]) == set(a.option_strings), (script, flag, a.option_strings)
        assert (a.type.__name__ if a.type else None) == flag["type"], (script, flag, a.type)
        assert a.default == flag["default"], (script, flag, a.default)
    # the additions never shadow a reference flag
    assert mod.parser.parse_args([]) is not None
