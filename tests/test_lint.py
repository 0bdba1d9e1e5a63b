"""CPU test: every name a function reads resolves to a local, an enclosing-scope, a module-level or a builtin name.  Most of the package only
runs on a GPU (no CPU path), so a misspelt variable in a CUDA-only branch would otherwise first show up on the GPU box."""
import builtins
import glob
import os
import symtable

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def unresolved_names(path):
  top = symtable.symtable(open(path).read(), path, "exec")
  module = {s.get_name() for s in top.get_symbols() if s.is_assigned() or s.is_imported() or s.is_namespace() or s.is_parameter()}

  def declared(tab):
    for s in tab.get_symbols():
      if s.is_declared_global() and s.is_assigned():
        module.add(s.get_name())
    for c in tab.get_children():
      declared(c)
  declared(top)
  bad = []

  def walk(tab):
    for s in tab.get_symbols():
      n = s.get_name()
      if s.is_referenced() and s.is_global() and n not in module and not hasattr(builtins, n) and n not in ("__file__", "__name__", "__doc__"):
        bad.append((tab.get_name(), n))
    for c in tab.get_children():
      walk(c)
  walk(top)
  return bad


def test_checker_sees_an_undefined_name(tmp_path):
  p = tmp_path / "x.py"
  p.write_text("import os\ndef f(a):\n  b = a + 1\n  return undefined_thing(b) + os.sep\nclass K:\n  def m(self):\n    return othername\n")
  assert sorted(n for _, n in unresolved_names(str(p))) == ["othername", "undefined_thing"]


def test_no_unresolved_names_in_the_repo():
  files = glob.glob(os.path.join(ROOT, "fasterrcnn_b200", "**", "*.py"), recursive = True) + [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]
  files += glob.glob(os.path.join(ROOT, "tools", "*.py")) + glob.glob(os.path.join(ROOT, "tests", "*.py")) + glob.glob(os.path.join(ROOT, "oracle", "*.py"))
  assert len(files) > 30
  bad = [(os.path.relpath(f, ROOT),) + b for f in files for b in unresolved_names(f)]
  assert not bad, bad
