"""Minimal stand-in for ``gin-config`` (pinned 0.3.0 by the reference, environment.yml:21).

gin is not installed in this image and there is no network, but the reference's
configuration is *semantic*: ``augment.simclr()`` instantiates its layers with no
arguments (augment/__init__.py:106-112) and relies on gin bindings keyed by class
name (configs/defaults/augment.gin:1-20).  This module implements exactly the subset
of gin the reference uses:

* ``@gin.configurable`` / ``@gin.configurable("name")`` / ``@gin.configurable(whitelist=[...])``
  on functions and classes (class ``__init__`` is wrapped in place, so ``isinstance``
  and ``nn.Module`` registration are unaffected);
* ``gin.REQUIRED``;
* ``gin.parse_config_files_and_bindings(files, bindings)``, ``parse_config``,
  ``bind_parameter``, ``query_parameter``, ``clear_config``, ``config_str``;
* the grammar found in ``configs/**/*.gin``: ``Name.param = <python literal>`` and
  ``# comments`` (statements may span lines inside brackets).

It is installed under the name ``gin`` by :func:`contrad_b200.dropin.install` only when
the real package is absent.
"""
import ast
import functools
import inspect
import threading

__all__ = [
    "REQUIRED", "configurable", "external_configurable", "parse_config", "parse_config_file",
    "parse_config_files_and_bindings", "bind_parameter", "query_parameter", "clear_config",
    "config_str", "operative_config_str",
]


class _Required(object):
    def __repr__(self):
        return "gin.REQUIRED"


REQUIRED = _Required()

_LOCK = threading.RLock()
_BINDINGS = {}      # configurable name -> {param: value}
_REGISTRY = {}      # configurable name -> (callable, whitelist, blacklist)
_OPERATIVE = {}     # bindings that were actually injected


def clear_config():
    with _LOCK:
        _BINDINGS.clear()
        _OPERATIVE.clear()


def bind_parameter(key, value):
    name, _, param = key.rpartition(".")
    if not name:
        raise ValueError("binding key must look like 'Name.param', got %r" % (key,))
    name = name.split("/")[-1]          # scopes are not used by the reference
    with _LOCK:
        _BINDINGS.setdefault(name, {})[param] = value


def query_parameter(key):
    name, _, param = key.rpartition(".")
    name = name.split("/")[-1]
    with _LOCK:
        try:
            return _BINDINGS[name][param]
        except KeyError:
            raise ValueError("no binding for %r" % (key,))


def _logical_lines(text):
    """Yield statements; a statement continues while brackets are open."""
    buf, depth = "", 0
    for raw in text.splitlines():
        line, in_str, out = raw, None, []
        for ch in line:
            if in_str:
                out.append(ch)
                if ch == in_str:
                    in_str = None
                continue
            if ch in "\"'":
                in_str = ch
            elif ch == "#":
                break
            out.append(ch)
        line = "".join(out).rstrip()
        if not line.strip() and depth == 0:
            continue
        depth += sum(line.count(c) for c in "([{") - sum(line.count(c) for c in ")]}")
        buf = (buf + " " + line.strip()) if buf else line.strip()
        if depth <= 0:
            yield buf
            buf, depth = "", 0
    if buf:
        yield buf


def parse_config(config, skip_unknown=False):
    if isinstance(config, (list, tuple)):
        config = "\n".join(config)
    for stmt in _logical_lines(config):
        if "=" not in stmt:
            raise SyntaxError("unsupported gin statement: %r" % (stmt,))
        key, _, value = stmt.partition("=")
        key, value = key.strip(), value.strip()
        try:
            parsed = ast.literal_eval(value)
        except (ValueError, SyntaxError):
            raise SyntaxError("only python literals are supported on the right-hand side: %r" % (stmt,))
        bind_parameter(key, parsed)


def parse_config_file(path, skip_unknown=False):
    with open(path, "r") as f:
        parse_config(f.read(), skip_unknown=skip_unknown)


def parse_config_files_and_bindings(config_files, bindings, finalize_config=True, skip_unknown=False):
    for path in (config_files or []):
        parse_config_file(path, skip_unknown=skip_unknown)
    if bindings:
        parse_config(bindings, skip_unknown=skip_unknown)


def config_str():
    with _LOCK:
        lines = []
        for name in sorted(_BINDINGS):
            for param in sorted(_BINDINGS[name]):
                lines.append("%s.%s = %r" % (name, param, _BINDINGS[name][param]))
        return "\n".join(lines) + ("\n" if lines else "")


def operative_config_str():
    with _LOCK:
        lines = []
        for name in sorted(_OPERATIVE):
            for param in sorted(_OPERATIVE[name]):
                lines.append("%s.%s = %r" % (name, param, _OPERATIVE[name][param]))
        return "\n".join(lines) + ("\n" if lines else "")


def _make_wrapper(fn, name, whitelist, blacklist, skip_first):
    sig = inspect.signature(fn)
    params = list(sig.parameters.values())
    if skip_first:
        params = params[1:]
    pos_names = [p.name for p in params
                 if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]
    accepts_kwargs = any(p.kind == p.VAR_KEYWORD for p in params)
    known = {p.name for p in params if p.kind not in (p.VAR_POSITIONAL, p.VAR_KEYWORD)}
    required_defaults = {p.name for p in params if p.default is REQUIRED}

    def allowed(param):
        if whitelist is not None and param not in whitelist:
            return False
        if blacklist is not None and param in blacklist:
            return False
        return accepts_kwargs or param in known

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        nargs = len(args) - (1 if skip_first else 0)
        given = set(pos_names[:nargs]) | set(kwargs)
        with _LOCK:
            bound = dict(_BINDINGS.get(name, {}))
        for param, value in bound.items():
            if param in given:
                continue
            if not allowed(param):
                raise ValueError("configurable %r has no configurable parameter %r" % (name, param))
            kwargs[param] = value
            with _LOCK:
                _OPERATIVE.setdefault(name, {})[param] = value
        missing = [p for p in required_defaults if p not in given and p not in kwargs]
        if missing:
            raise RuntimeError("required bindings for %r not provided in config: %s"
                               % (name, sorted(missing)))
        return fn(*args, **kwargs)

    return wrapper


def _decorate(target, name, whitelist, blacklist):
    name = name or target.__name__
    if inspect.isclass(target):
        target.__init__ = _make_wrapper(target.__init__, name, whitelist, blacklist, skip_first=True)
        decorated = target
    else:
        decorated = _make_wrapper(target, name, whitelist, blacklist, skip_first=False)
    with _LOCK:
        _REGISTRY[name] = (decorated, whitelist, blacklist)
    return decorated


def configurable(name_or_fn=None, module=None, whitelist=None, blacklist=None, allowlist=None, denylist=None):
    whitelist = whitelist if whitelist is not None else allowlist
    blacklist = blacklist if blacklist is not None else denylist
    if callable(name_or_fn) and not isinstance(name_or_fn, str):
        return _decorate(name_or_fn, None, whitelist, blacklist)

    def deco(target):
        return _decorate(target, name_or_fn, whitelist, blacklist)
    return deco


def external_configurable(fn, name=None, module=None, whitelist=None, blacklist=None):
    return _decorate(fn, name, whitelist, blacklist)
