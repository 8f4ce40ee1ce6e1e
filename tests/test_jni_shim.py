"""jni/b200shuffle_jni.c — the JNI layer a maintainer builds on the reference side.  No JDK exists
here, so the file is compiled against jni/stub/jni.h (types + the JNIEnv members the shim uses) with -Wall -Wextra
-Werror: argument order and types against include/b200shuffle.h, the exported Java_* names, and the pairing of
Get/ReleasePrimitiveArrayCritical are compiler- and script-checked.  It cannot be RUN without a JVM."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "jni", "b200shuffle_jni.c")


def test_shim_compiles_against_the_c_abi_and_exports_every_native(tmp_path):
    import spark_s3_shuffle_b200 as pkg

    lib = pkg.build()
    so = str(tmp_path / "libb200shuffle_jni.so")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-fPIC", "-shared",
                           "-I" + os.path.join(ROOT, "jni", "stub"), "-I" + os.path.join(ROOT, "include"), SRC, "-o", so,
                           "-L" + os.path.dirname(lib), "-l:libb200shuffle.so", "-Wl,-z,defs"])
    exported = set(re.findall(r" T Java_org_apache_spark_shuffle_gpu_B200Codec_(\w+)",
                              subprocess.run(["nm", "-D", so], capture_output=True, text=True).stdout))
    declared = set(re.findall(r"J\((\w+)\)\(JNIEnv", open(SRC).read()))
    assert exported == declared and len(exported) >= 19
    # every @native of the Scala object in INTEGRATION.md has its C function, and vice versa
    natives = set(re.findall(r"@native def (\w+)\(", open(os.path.join(ROOT, "INTEGRATION.md")).read()))
    assert natives == exported, (natives ^ exported)


def test_every_critical_get_is_released_on_every_path():
    src = open(SRC).read()
    bodies = re.split(r"\nJNIEXPORT ", src)[1:]
    assert len(bodies) >= 19
    for b in bodies:
        name = re.search(r"J\((\w+)\)", b).group(1)
        gets = re.findall(r"(\w+) = crit_get\(e, (\w+)\)", b)
        puts = re.findall(r"crit_put\(e, (\w+), (\w+), (0|JNI_ABORT)\)", b)
        assert sorted((a, v) for v, a in gets) == sorted((a, v) for a, v, _ in puts), name
        if gets:  # single exit after the releases: no early return between the first get and the last put
            first, last = b.index("crit_get"), b.rindex("crit_put")
            assert "return" not in b[first:last], name
