"""poor man's pyflakes: report names that are loaded in a function but bound nowhere (module, enclosing functions, builtins)"""
import ast, builtins, sys

def check(path):
    tree = ast.parse(open(path).read())
    mod_names = set(dir(builtins))
    for n in ast.walk(tree):
        if isinstance(n, (ast.Import, ast.ImportFrom)):
            for a in n.names:
                mod_names.add((a.asname or a.name).split('.')[0])
        elif isinstance(n, (ast.FunctionDef, ast.ClassDef, ast.AsyncFunctionDef)):
            mod_names.add(n.name)
        elif isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
            pass
    for n in tree.body:
        for t in ast.walk(n) if isinstance(n, (ast.Assign, ast.AugAssign, ast.AnnAssign, ast.For, ast.With, ast.Try, ast.If)) else []:
            if isinstance(t, ast.Name) and isinstance(t.ctx, ast.Store):
                mod_names.add(t.id)
    bad = []
    def visit(fn, outer):
        bound = set(outer)
        args = fn.args
        for a in args.args + args.kwonlyargs + args.posonlyargs + ([args.vararg] if args.vararg else []) + ([args.kwarg] if args.kwarg else []):
            bound.add(a.arg)
        for t in ast.walk(fn):
            if isinstance(t, ast.Name) and isinstance(t.ctx, ast.Store):
                bound.add(t.id)
            elif isinstance(t, (ast.FunctionDef, ast.ClassDef)) and t is not fn:
                bound.add(t.name)
            elif isinstance(t, (ast.Import, ast.ImportFrom)):
                for a in t.names:
                    bound.add((a.asname or a.name).split('.')[0])
            elif isinstance(t, ast.ExceptHandler) and t.name:
                bound.add(t.name)
            elif isinstance(t, (ast.Global, ast.Nonlocal)):
                bound.update(t.names)
            elif isinstance(t, ast.arg):
                bound.add(t.arg)
        for t in ast.walk(fn):
            if isinstance(t, ast.Name) and isinstance(t.ctx, ast.Load) and t.id not in bound:
                bad.append((path, t.lineno, t.id))
    def walk(node, outer):
        for ch in ast.iter_child_nodes(node):
            if isinstance(ch, (ast.FunctionDef, ast.AsyncFunctionDef)):
                visit(ch, outer)
                inner = set(outer)
                for t in ast.walk(ch):
                    if isinstance(t, ast.Name) and isinstance(t.ctx, ast.Store): inner.add(t.id)
                    if isinstance(t, ast.arg): inner.add(t.arg)
                walk(ch, inner)
            else:
                walk(ch, outer)
    walk(tree, mod_names)
    return sorted(set(bad))

rc = 0
for p in sys.argv[1:]:
    for b in check(p):
        print('%s:%d: undefined name %s' % b); rc = 1
sys.exit(rc)
