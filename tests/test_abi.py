"""CPU-side checks of the drop-in boundary: the library builds, loads and exports every declared symbol."""
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol(sceneprep_lib):
    from garden_b200.binding import SYMBOLS, load_library
    lib = load_library()
    header = (ROOT / "include" / "garden_sceneprep.h").read_text()
    declared = set(re.findall(r"\b(gsp_[a-z_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} is declared in include/garden_sceneprep.h but not exported"
    assert declared == set(SYMBOLS), f"binding and header disagree: {declared ^ set(SYMBOLS)}"
    assert b"sm_100a" in lib.gsp_version()
