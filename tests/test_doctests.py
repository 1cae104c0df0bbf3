"""Runs every doctest of the package (builders' examples etc.), as the reference does with ``--doctest-modules``
(``/root/reference/pytest.ini:2``) -- here as an ordinary test so that ``pytest tests/`` covers them too."""
import doctest
import importlib
import pkgutil

import myrtlespeech_b200


def test_package_doctests():
    attempted = 0
    for info in pkgutil.walk_packages(myrtlespeech_b200.__path__, prefix="myrtlespeech_b200."):
        mod = importlib.import_module(info.name)
        res = doctest.testmod(mod, optionflags=doctest.ELLIPSIS | doctest.NORMALIZE_WHITESPACE)
        assert res.failed == 0, f"{info.name}: {res.failed} doctest(s) failed"
        attempted += res.attempted
    assert attempted >= 8, attempted     # the three builder examples (several statements each)
