"""CPU tests: the C-ABI library loads and exports every symbol include/frcnn_b200.h declares."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
  text = open(os.path.join(ROOT, "include", "frcnn_b200.h")).read()
  text = re.sub(r"/\*.*?\*/", "", text, flags = re.S)
  return sorted(set(re.findall(r"\b(frcnn_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
  from fasterrcnn_b200 import _lib
  if not os.path.exists(_lib.LIB_PATH):
    import __graft_entry__ as g
    g.build()
  handle = _lib.lib()
  declared = _declared_symbols()
  assert len(declared) >= 30
  for name in declared:
    assert hasattr(handle, name), "missing export: " + name
  assert sorted(_lib.exported_symbols()) == declared, set(_lib.exported_symbols()) ^ set(declared)
  assert handle.frcnn_version() >= 100


def test_no_cpu_fallback():
  """Ops refuse CPU tensors instead of silently computing somewhere else."""
  import torch as t
  from fasterrcnn_b200 import ops, _lib
  with pytest.raises(_lib.FrcnnError):
    ops.conv2d_act(t.zeros((1, 4, 8, 8)), t.zeros((4, 4, 3, 3)), t.zeros((4,)))
  with pytest.raises(_lib.FrcnnError):
    ops.nms(t.zeros((4, 4)), t.zeros((4,)), 0.5)


def test_product_never_imports_oracle():
  pkg = os.path.join(ROOT, "fasterrcnn_b200")
  for dirpath, _, files in os.walk(pkg):
    for f in files:
      if f.endswith((".py", ".cu", ".cuh", ".h")):
        src = open(os.path.join(dirpath, f)).read()
        assert "oracle" not in src.replace("# oracle", ""), "%s mentions the oracle" % f


def test_state_dict_keys_match_reference_layout():
  import fasterrcnn_b200 as f
  from oracle import frcnn_oracle as orc
  model = f.FasterRCNNModel(21, f.vgg16.VGG16Backbone(0.0))
  sd = model.state_dict()
  shapes = orc.vgg16_param_shapes()
  assert list(sd.keys()) == list(shapes.keys())
  for k, shape in shapes.items():
    assert tuple(sd[k].shape) == tuple(shape)
  frozen = [k for k, p in model.named_parameters() if not p.requires_grad]
  assert sorted(frozen) == sorted(k for k in shapes if k not in orc.trainable_keys_vgg16(shapes))


def test_sass_carries_the_blackwell_instructions():
  """The built library really is sm_100a tcgen05 / TMA / PDL code (B200_PROFILING.md: the SASS mnemonics that prove it): every kernel
  entry has griddepcontrol (PREEXIT + ACQBULK), the GEMM kernel has UTCHMMA (tcgen05.mma), UTMALDG (TMA tile loads), LDTM (tcgen05.ld)."""
  import shutil
  import subprocess
  from fasterrcnn_b200 import _lib
  cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
  if not os.path.exists(cuobjdump):
    pytest.skip("cuobjdump not available")
  sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output = True, text = True, timeout = 300).stdout
  assert "sm_100a" in sass
  kernels = len(re.findall(r"^\s*Function : ", sass, flags = re.M))
  assert kernels >= 50
  count = lambda op: len(re.findall(r"\b%s\b" % op, sass))
  assert count("UTCHMMA") > 0 and count("UTMALDG") > 0 and count("LDTM") > 0
  pair_kernels = len(re.findall(r"^\s*Function : \S*tc_conv_kernelILi\dELi\d+ELi\dELb1ELb1E", sass, flags = re.M))
  assert count("ACQBULK") == kernels                                             # one griddepcontrol.wait per kernel
  assert count("PREEXIT") == kernels + pair_kernels                              # one launch_dependents per kernel; the pair kernels carry an early AND a late one (one executes)
  # determinism: the only atomics are integer ones (order-independent); no floating-point RED / ATOM anywhere
  assert not re.search(r"\b(RED|REDG|ATOM|ATOMG|ATOMS)\.[A-Z0-9.]*F(16|32|64)", sass)
  # PDL safety, statically: in every kernel nothing touches global memory before griddepcontrol.wait (ACQBULK).  ptxas is free to move
  # non-coherent loads (ld.global.nc, what `const __restrict__` turns into) above the wait -- it did so once for a device-side count --
  # so the order is checked on the SASS of every build.  The tcgen05 kernel reads its TMEM slot from shared memory through a generic
  # pointer before the wait (LD.E), nothing else; a branch before the wait may only jump within the pre-wait region.
  # (opcodes, anchored at the start of the instruction behind an optional predicate: `SYNCS.ARRIVE.TRANS64.RED` is a shared-memory mbarrier op)
  strict = re.compile(r"^(@!?U?P\d+\s+)?(LDG|STG|ATOMG|REDG|ATOM|RED|LDGSTS|UTMALDG|UTMASTG|LDGMC|UBLKCP)\b")
  generic = re.compile(r"^(@!?U?P\d+\s+)?(LD|ST)(\.E)?\b")
  for body in re.split(r"^\s*Function : ", sass, flags = re.M)[1:]:
    name = body.split("\n", 1)[0].strip()
    ins = [(int(m.group(1), 16), re.sub(r"/\*.*?\*/", "", line).strip()) for line in body.split("\n") for m in [re.search(r"/\*([0-9a-f]{4,})\*/", line)] if m]
    at = next(i for i, (_, op) in enumerate(ins) if "ACQBULK" in op)
    wait_addr = ins[at][0]
    for addr, op in ins[:at]:
      assert not strict.search(op), (name, op)
      assert "tc_conv_kernel" in name or not generic.search(op), (name, op)
      m = re.search(r"\bBRA\s+(0x[0-9a-f]+)", op)
      if m is not None and int(m.group(1), 16) > wait_addr:
        # an out-of-line block (ptxas moves the retry loops of tcgen05.alloc.cta_group::2 behind the kernel body): it must come back
        # to the pre-wait region by an unconditional branch without touching global memory on the way
        by_addr = {a: o for a, o in ins}
        pc, steps = int(m.group(1), 16), 0
        while True:
          o = by_addr[pc]
          assert not strict.search(o) and not generic.search(o), (name, op, o)
          back = re.match(r"^BRA\s+(0x[0-9a-f]+)", o)
          if back is not None:
            assert int(back.group(1), 16) <= wait_addr, (name, op, o)
            break
          pc += 16; steps += 1
          assert steps < 64, (name, op)


def test_sm_reserve_changes_neither_workspace_sizes_nor_amax_slots():
  """frcnn_set_sm_reserve only shrinks the CTA count of the persistent GEMM launches: the sizes callers cache per geometry (workspace
  bytes, number of per-CTA output maxima) must not depend on it.  Pure host calls -- no device needed."""
  from fasterrcnn_b200 import _lib
  L = _lib.lib()
  geoms = [(1, 600, 1000, 64, 64, 3, 3, 1, 1), (1, 150, 250, 256, 256, 3, 3, 1, 1), (1, 37, 62, 512, 512, 3, 3, 1, 1), (1, 38, 63, 1024, 256, 1, 1, 1, 0),
           (128, 1, 1, 25088, 4096, 1, 1, 1, 0), (128, 1, 1, 4096, 4096, 1, 1, 1, 0), (2, 75, 125, 512, 512, 3, 3, 1, 1)]
  def sizes():
    out = []
    for g in geoms:
      for eng in (_lib.ENGINE_TC_3XF16, _lib.ENGINE_AUTO):
        out.append((L.frcnn_conv2d_fwd_workspace_bytes(*g, eng), L.frcnn_conv2d_dgrad_workspace_bytes(*g, eng), L.frcnn_conv2d_wgrad_workspace_bytes(*g, eng)))
      out.append((L.frcnn_conv2d_amax_slots(0, *g), L.frcnn_conv2d_amax_slots(1, *g)))
    return out
  base = sizes()
  assert any(s[0] > 0 for s in base)
  try:
    for reserve in (8, 20, 64, 1000):
      _lib.set_sm_reserve(reserve)
      assert sizes() == base, reserve
  finally:
    assert _lib.set_sm_reserve(0) == 132          # 1000 was clamped to 148 - 16
  assert max(s[0] for s in base[2::3]) <= 148     # amax slots = CTAs of an unreserved launch


def test_ctypes_signatures_match_the_header_prototypes():
  """Every binding in fasterrcnn_b200/_lib.py has the argument list of its prototype in include/frcnn_b200.h: same count, and per argument
  pointer -> c_void_p, int -> c_int, float -> c_float, double -> c_double, size_t -> c_size_t (a mismatch would only show on the GPU box,
  as a corrupted call)."""
  import ctypes
  from fasterrcnn_b200 import _lib
  text = open(os.path.join(ROOT, "include", "frcnn_b200.h")).read()
  text = re.sub(r"/\*.*?\*/", "", text, flags = re.S)
  protos = dict((m.group(2), (m.group(1).strip(), m.group(3))) for m in re.finditer(r"^([A-Za-z_][\w \*]*?)\b(frcnn_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags = re.M | re.S))
  assert set(protos) == set(_lib._SIGNATURES), set(protos) ^ set(_lib._SIGNATURES)

  def ctype_of(param):
    param = " ".join(param.split())
    if "*" in param:
      return ctypes.c_void_p
    base = param.rsplit(" ", 1)[0] if " " in param else param
    base = base.replace("const ", "").strip()
    return {"int": ctypes.c_int, "float": ctypes.c_float, "double": ctypes.c_double, "size_t": ctypes.c_size_t, "int32_t": ctypes.c_int}[base]

  for name, (ret, params) in protos.items():
    restype, argtypes = _lib._SIGNATURES[name]
    plist = [] if params.strip() in ("", "void") else [p for p in params.split(",")]
    assert len(plist) == len(argtypes), (name, len(plist), len(argtypes))
    for i, (p, a) in enumerate(zip(plist, argtypes)):
      assert ctype_of(p) is a, (name, i, p.strip(), a)
    want = ctypes.c_char_p if ("char" in ret and "*" in ret) else (None if ret == "void" else {"int": ctypes.c_int, "size_t": ctypes.c_size_t}[ret.replace("const ", "").strip()])
    assert restype is want, (name, ret, restype)


def test_call_sites_pass_the_bound_number_of_arguments():
  """Static count of the arguments at every `...frcnn_xxx(...)` call in the package against the binding (`*geom` = the 9 geometry
  integers): ctypes would raise only when the call is reached -- on the GPU box."""
  import ast
  import glob
  from fasterrcnn_b200 import _lib
  checked = 0
  for path in glob.glob(os.path.join(ROOT, "fasterrcnn_b200", "*.py")) + [os.path.join(ROOT, "bench.py")] + glob.glob(os.path.join(ROOT, "tools", "*.py")):
    tree = ast.parse(open(path).read())
    for node in ast.walk(tree):
      if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr in _lib._SIGNATURES:
        n, known = 0, True
        for a in node.args:
          if isinstance(a, ast.Starred):
            if isinstance(a.value, ast.Name) and a.value.id in ("geom", "g9"):
              n += 9
            else:
              known = False
          else:
            n += 1
        if known and not node.keywords:
          assert n == len(_lib._SIGNATURES[node.func.attr][1]), (os.path.basename(path), node.lineno, node.func.attr, n, len(_lib._SIGNATURES[node.func.attr][1]))
          checked += 1
  assert checked >= 60


def test_python_constants_match_the_header_defines():
  from fasterrcnn_b200 import _lib
  text = open(os.path.join(ROOT, "include", "frcnn_b200.h")).read()
  defs = {m.group(1): int(m.group(2).strip("()")) for m in re.finditer(r"^#define (FRCNN_[A-Z0-9_]+) (\(?-?\d+\)?)", text, flags = re.M)}
  assert (defs["FRCNN_ACT_NONE"], defs["FRCNN_ACT_RELU"], defs["FRCNN_ACT_SIGMOID"]) == (_lib.ACT_NONE, _lib.ACT_RELU, _lib.ACT_SIGMOID)
  assert (defs["FRCNN_ENGINE_AUTO"], defs["FRCNN_ENGINE_SIMT_FP32"], defs["FRCNN_ENGINE_TC_3XTF32"], defs["FRCNN_ENGINE_TC_3XF16"]) == \
         (_lib.ENGINE_AUTO, _lib.ENGINE_SIMT_FP32, _lib.ENGINE_TC_3XTF32, _lib.ENGINE_TC_3XF16)
  assert defs["FRCNN_OK"] == 0 and defs["FRCNN_E_BADARG"] < 0
