"""The XLA-FFI side of the boundary (integration/) cannot be built or imported in this image (no jaxlib / ksim).  What CAN be
checked without them, and is checked here:

  * integration/zb_xla_ffi.cc type-checks (g++ -fsyntax-only) against the stand-in FFI header integration/stub/, whose
    XLA_FFI_DEFINE_HANDLER_SYMBOL static_asserts each handler against the types of its Bind() chain;
  * for every handler, the operand / result / attribute lists of its C++ signature, of its Bind() chain and of the Python tables in
    integration/zbot_cuda_engine.py agree in count, order, name and dtype;
  * ZbotData carries every mjx.Data field the reference task reads, and the engine class defines ksim's reset / step signatures.
"""
import ast
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CC = os.path.join(ROOT, "integration", "zb_xla_ffi.cc")
PY = os.path.join(ROOT, "integration", "zbot_cuda_engine.py")


def test_ffi_source_typechecks_against_stub_header():
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    inc = ["-I" + os.path.join(ROOT, "integration", "stub"), "-I" + os.path.join(ROOT, "include")]
    cuda_inc = "/usr/local/cuda/include"
    if not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("no CUDA headers")
    r = subprocess.run([gxx, "-fsyntax-only", "-std=c++17", "-Wall", *inc, "-I" + cuda_inc, CC], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    # the stub's static_assert really bites: dropping one operand from a Bind() chain must fail to compile
    src = open(CC).read()
    broken = src.replace(".Arg<F32>().Arg<F32>().Arg<U8>().Arg<F32>().Ret<F32>().Ret<F32>()", ".Arg<F32>().Arg<U8>().Arg<F32>().Ret<F32>().Ret<F32>()")
    assert broken != src
    r2 = subprocess.run([gxx, "-fsyntax-only", "-std=c++17", *inc, "-I" + cuda_inc, "-x", "c++", "-"], input=broken, capture_output=True, text=True)
    assert r2.returncode != 0 and "not callable with" in r2.stderr


def _py_tables():
    tree = ast.parse(open(PY).read())
    env = {"NQ": 27, "NV": 26, "NU": 20, "NB": 26, "NSITE": 3}
    out = {}
    for node in tree.body:
        if isinstance(node, ast.Assign) and isinstance(node.targets[0], ast.Name) and node.targets[0].id.startswith("_") and node.targets[0].id.isupper() is False:
            pass
        if isinstance(node, ast.Assign) and isinstance(node.targets[0], ast.Name):
            name = node.targets[0].id
            if re.fullmatch(r"_[A-Z_]+", name) and name != "_DT":
                out[name] = eval(compile(ast.Expression(node.value), PY, "eval"), {}, {**env, **out})  # noqa: S307 (our own file)
    return tree, out


def _cc_handlers():
    src = open(CC).read()
    handlers = {}
    for m in re.finditer(r"static ffi::Error (\w+)\(\s*(.*?)\)\s*\{", src, re.S):
        sig = re.sub(r"//[^\n]*", "", m.group(2))
        params = [p.strip().rsplit(" ", 1) for p in sig.replace("\n", " ").split(",") if p.strip()]
        handlers[m.group(1)] = [(t.strip(), n.strip()) for t, n in params]
    binds = {}
    for m in re.finditer(r"XLA_FFI_DEFINE_HANDLER_SYMBOL\(\s*(\w+),\s*(\w+),(.*?)\);", src, re.S):
        chain = re.sub(r"//[^\n]*", "", m.group(3))
        binds[m.group(2)] = (m.group(1), re.findall(r"\.(Ctx|Attr|Arg|Ret)<([^>]*(?:<[^>]*>)?)>\((?:\"(\w+)\")?\)", chain))
    return handlers, binds


CT = {"f32": "F32", "u8": "U8", "i32": "S32"}
ATTR = {"i64": "int64_t", "i32": "int32_t", "f32": "float"}


@pytest.mark.parametrize("impl,symbol,attrs,operands,results", [
    ("ZbStepObsImpl", "zb_step_obs_ffi", "_STEP_ATTRS", ("_STATE", "_STEP_INPUTS", "_MODEL_FIELDS"), "_STEP_RESULTS"),
    ("ZbResetObsImpl", "zb_reset_obs_ffi", "_RESET_ATTRS", ("_RESET_INPUTS", "_MODEL_FIELDS"), "_RESET_RESULTS"),
    ("ZbActorImpl", "zb_actor_ffi", "_ACTOR_ATTRS", ("_ACTOR_WEIGHTS", "_ACTOR_INPUTS"), "_ACTOR_RESULTS"),
])
def test_operand_order_matches_between_cpp_and_python(impl, symbol, attrs, operands, results):
    _, t = _py_tables()
    handlers, binds = _cc_handlers()
    sig = handlers[impl]
    sym, chain = binds[impl]
    assert sym == symbol
    want_attrs = list(t[attrs])
    want_ops = [x for name in operands for x in t[name]]
    want_res = list(t[results])
    # C++ signature: stream, attributes, operands, results — names and types in table order
    assert sig[0] == ("cudaStream_t", "stream")
    i = 1
    for name, dt in want_attrs:
        assert sig[i] == (ATTR[dt], name), (sig[i], name)
        i += 1
    for name, _, dt in want_ops:
        assert sig[i] == (CT[dt], name), (sig[i], name)
        i += 1
    state_names = {n for n, _, _ in t["_STATE"]} | {"carry"}
    for name, _, dt in want_res:
        cname = name + "_o" if name in state_names else name  # state results are the `_o` twins of the aliased operands
        assert sig[i] == ("R" + CT[dt], cname), (sig[i], cname)
        i += 1
    assert i == len(sig)
    # Bind() chain: same kinds, types and attribute names in the same order
    kinds = [c[0] for c in chain]
    assert kinds == ["Ctx"] + ["Attr"] * len(want_attrs) + ["Arg"] * len(want_ops) + ["Ret"] * len(want_res)
    assert [c[2] for c in chain if c[0] == "Attr"] == [n for n, _ in want_attrs]
    assert [c[1] for c in chain if c[0] == "Attr"] == [ATTR[dt] for _, dt in want_attrs]
    assert [c[1] for c in chain if c[0] == "Arg"] == [CT[dt] for _, _, dt in want_ops]
    assert [c[1] for c in chain if c[0] == "Ret"] == [CT[dt] for _, _, dt in want_res]


def test_engine_class_shape_and_data_fields():
    tree, t = _py_tables()
    classes = {n.name: n for n in tree.body if isinstance(n, ast.ClassDef)}
    eng = classes["ZbotCudaEngine"]
    assert [ast.unparse(b) for b in eng.bases] == ["ksim.PhysicsEngine"]
    methods = {n.name: [a.arg for a in n.args.args] for n in eng.body if isinstance(n, ast.FunctionDef)}
    assert methods["reset"] == ["self", "physics_model", "curriculum_level", "rng"]                      # ksim PhysicsEngine.reset [R]
    assert methods["step"] == ["self", "action", "physics_model", "physics_state", "curriculum_level", "rng"]  # PhysicsEngine.step [R]
    fields = [n.target.id for n in classes["ZbotData"].body if isinstance(n, ast.AnnAssign)]
    # every mjx.Data field the task reads
    for f in ("qpos", "qvel", "qacc", "xpos", "xquat", "cinert", "cvel", "site_xpos", "site_xmat", "sensordata", "actuator_force", "ctrl", "time"):
        assert f in fields, f
    produced = {n for n, _, _ in t["_STEP_RESULTS"]}
    renamed = {"qpos", "qvel", "qacc_warmstart", "time", "imu_x", "imu_lag"}
    assert set(fields) <= produced | renamed | {"done"}
    # shapes the kernel writes (include/zbphys.h): spot checks against the header's compile-time sizes
    hdr = open(os.path.join(ROOT, "include", "zbphys.h")).read()
    sizes = {k: int(v) for k, v in re.findall(r"#define (ZB_N\w+) (\d+)", hdr)}
    shapes = {n: s for n, s, _ in t["_STEP_RESULTS"]}
    assert shapes["actor_obs"] == (sizes["ZB_NACTOR_OBS"],) and shapes["critic_obs"] == (sizes["ZB_NCRITIC_OBS"],)
    assert shapes["sensordata"] == (sizes["ZB_NSENSORDATA"],) and shapes["reward_terms"] == (sizes["ZB_NREWARD_TERMS"],)
    assert shapes["xpos"] == (sizes["ZB_NBODY"], 3) and shapes["qpos"] == (sizes["ZB_NQ"],)
