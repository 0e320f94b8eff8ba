"""The oracle is test infrastructure: nothing under mockingbird_b200/ (the product) may import, load or execute anything under oracle/,
and the product has no CPU fallback - importing the package on a machine without CUDA must not silently route anywhere else."""
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
ORACLE_MODULES = sorted(p.stem for p in (ROOT / "oracle").glob("*.py"))


def test_product_sources_never_reference_the_oracle():
    assert ORACLE_MODULES, "oracle/ has no modules?"
    pat = re.compile(r"^\s*(?:from|import)\s+(" + "|".join(map(re.escape, ORACLE_MODULES)) + r")\b", re.M)
    offenders = []
    for p in (ROOT / "mockingbird_b200").rglob("*.py"):
        txt = p.read_text(errors="ignore")
        if pat.search(txt) or re.search(r"[\"'/]oracle[\"'/]", txt):
            offenders.append(str(p.relative_to(ROOT)))
    for p in (ROOT / "mockingbird_b200" / "csrc").glob("*"):
        if p.is_file() and p.suffix in (".cu", ".cpp", ".h", ".cuh") and re.search(r"#\s*include[^\n]*oracle", p.read_text(errors="ignore")):
            offenders.append(str(p.relative_to(ROOT)))  # (comments may name the twin; only an #include would link it in)
    assert not offenders, offenders


def test_importing_the_product_does_not_import_the_oracle():
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import mockingbird_b200\n"
        "from mockingbird_b200.vocoder.hifigan import inference as a\n"
        "from mockingbird_b200.vocoder.wavernn import inference as b\n"
        "from mockingbird_b200.synthesizer.inference import Synthesizer\n"
        "bad = [m for m in sys.modules if m in %r]\n"
        "print('LOADED', bad)\n" % (str(ROOT), ORACLE_MODULES)
    )
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "LOADED []" in r.stdout, r.stdout
