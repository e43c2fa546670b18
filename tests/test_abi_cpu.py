"""CPU: the C-ABI library loads, exports every symbol include/parasuite_b200.h declares, and refuses to
compute without a GPU (there is no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from parasuite_b200 import abi

HEADER = os.path.join(abi.INCLUDE_DIR, "parasuite_b200.h")


def declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ps_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(abi.EXPORTS)


def test_library_exports_every_declared_symbol():
    if not os.path.exists(abi.lib_path()):
        pytest.skip("library not built (run __graft_entry__.build())")
    lib = abi.load_library()          # raises AttributeError on a missing export
    assert lib.ps_abi_version() == abi.PS_ABI_VERSION
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.ps_profile_acc_len(51, 0) == 16 * 51 + 32 + 2 * 51 + abi.PS_PC_COUNT
    assert b"sorted" in lib.ps_strerror(abi.PS_ERR_UNSORTED)


def test_no_cpu_fallback():
    if not os.path.exists(abi.lib_path()):
        pytest.skip("library not built")
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = abi.load_library()
    h = C.c_void_p()
    assert lib.ps_create(C.byref(h), 0) == abi.PS_ERR_NO_DEVICE
    assert not h.value
    from parasuite_b200.runtime import Context
    with pytest.raises(abi.PsError):
        Context(0)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(abi, "_lib", None)
    monkeypatch.setattr(abi, "LIB_DIR", str(tmp_path))
    with pytest.raises(abi.NativeLibraryMissing):
        abi.load_library()


def test_jni_shim_type_checks_against_stub_header():
    """No JDK in this image: the JNI shim (jni/parasuite_jni.c) is compiled with -fsyntax-only against a stand-in jni.h
    that declares exactly the JNI calls it makes, together with the real C ABI header."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    repo = os.path.dirname(abi.INCLUDE_DIR)
    cmd = [gcc, "-std=c11", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I", os.path.join(repo, "tests", "jni_stub"),
           "-I", abi.INCLUDE_DIR, os.path.join(repo, "jni", "parasuite_jni.c")]
    subprocess.run(cmd, check=True, capture_output=True)
    src = open(os.path.join(repo, "jni", "parasuite_jni.c")).read()
    for sym in ("ps_create", "ps_destroy", "ps_reference_load_fasta", "ps_profile_bam", "ps_pileup_bam", "ps_pileup_next",
                "ps_pileup_counters_get", "ps_pileup_close"):
        assert sym in src and sym in abi.EXPORTS


def test_java_classes_declare_what_the_shim_exports():
    """jni/java/**/*.java (the classes a maintainer drops into the reference's tree) and jni/parasuite_jni.c agree on the
    set of native methods, class by class."""
    import glob
    import re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "jni")
    shim = open(os.path.join(root, "parasuite_jni.c")).read()
    exported = set(re.findall(r"JNICALL\s+Java_([A-Za-z0-9_]+)\(", shim))
    declared = set()
    for path in glob.glob(os.path.join(root, "java", "**", "*.java"), recursive=True):
        src = open(path).read()
        pkg = re.search(r"package\s+([\w.]+);", src).group(1)
        cls = re.search(r"public final class (\w+)", src).group(1)
        for name in re.findall(r"public static native [\w\[\]]+ (\w+)\(", src):
            declared.add(f"{pkg.replace('.', '_')}_{cls}_{name}")
    assert exported == declared and len(exported) >= 15
