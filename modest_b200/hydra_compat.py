"""A small stand-in for the parts of Hydra 1.x / OmegaConf the reference's CLIs use.

The reference decorates its three programs with `@hydra.main(config_path="configs/",
config_name=...)` (generate_cluster_mask/pre_compute_pp_score.py:83, generate_mask.py:30,
gen_label_files.py:31) and relies on: a `defaults:` list with one config group
(`- data_paths: fw70_2m.yaml`), `${key}` / `${hydra:runtime.cwd}` / `${hydra:run.dir}`
interpolation, `???` mandatory values, `key=value` command-line overrides (nested keys, lists,
`null`, group switches such as `data_paths=nusc.yaml`), the change of working directory into
`outputs/<date>/<time>`, attribute access on the config, `OmegaConf.to_yaml` and
`OmegaConf.save`.  Neither package is installed in the build image, so the drop-in CLIs use
this module when `import hydra` fails and the real thing when it succeeds
(`get_hydra()` below).
"""
from __future__ import annotations

import datetime
import functools
import inspect
import os
import re
import sys

import yaml


class MissingMandatoryValue(Exception):
    pass


class DictConfig(dict):
    """dict with attribute access; nested dicts are wrapped on the way out."""

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError as e:
            raise AttributeError(key) from e

    def __setattr__(self, key, value):
        self[key] = value

    def __getitem__(self, key):
        v = dict.__getitem__(self, key)
        if isinstance(v, str) and v == "???":
            raise MissingMandatoryValue(f"Missing mandatory value: {key}")
        return v

    def get(self, key, default=None):
        return self[key] if key in self else default


def _wrap(obj):
    if isinstance(obj, dict):
        return DictConfig({k: _wrap(v) for k, v in obj.items()})
    if isinstance(obj, (list, tuple)):
        return [_wrap(v) for v in obj]
    return obj


def _unwrap(obj):
    if isinstance(obj, dict):
        return {k: _unwrap(dict.__getitem__(obj, k)) for k in obj}
    if isinstance(obj, (list, tuple)):
        return [_unwrap(v) for v in obj]
    return obj


class OmegaConf:
    @staticmethod
    def to_yaml(cfg) -> str:
        return yaml.safe_dump(_unwrap(cfg), default_flow_style=False, sort_keys=False)

    @staticmethod
    def save(config, f):
        text = OmegaConf.to_yaml(config)
        if hasattr(f, "write"):
            f.write(text)
        else:
            with open(f, "w") as fh:
                fh.write(text)

    @staticmethod
    def create(obj=None):
        return _wrap(obj or {})

    @staticmethod
    def to_container(cfg, resolve=True):
        return _unwrap(cfg)


_INTERP = re.compile(r"\$\{([^${}]+)\}")


def _lookup(root, dotted):
    cur = root
    for part in dotted.split("."):
        cur = dict.__getitem__(cur, part) if isinstance(cur, dict) else cur[int(part)]
    return cur


def _resolve_value(val, root, hydra_vars, depth=0):
    if depth > 32:
        raise RecursionError("interpolation loop in config")
    if isinstance(val, dict):
        return {k: _resolve_value(v, root, hydra_vars, depth) for k, v in val.items()}
    if isinstance(val, list):
        return [_resolve_value(v, root, hydra_vars, depth) for v in val]
    if not isinstance(val, str) or "${" not in val:
        return val

    def one(expr):
        expr = expr.strip()
        if expr.startswith("hydra:"):
            return hydra_vars[expr[len("hydra:"):]]
        return _resolve_value(_lookup(root, expr), root, hydra_vars, depth + 1)

    whole = _INTERP.fullmatch(val)
    if whole:                                   # "${x}" keeps the referenced type
        return one(whole.group(1))
    prev = None
    while prev != val and "${" in val:
        prev = val
        val = _INTERP.sub(lambda m: str(one(m.group(1))), val)
    return val


def _set_dotted(cfg, dotted, value):
    parts = dotted.split(".")
    cur = cfg
    for p in parts[:-1]:
        if p not in cur or not isinstance(cur[p], dict):
            cur[p] = {}
        cur = cur[p]
    cur[parts[-1]] = value


def compose(config_dir, config_name, overrides=(), cwd=None, run_dir=None):
    """Load `config_name` from `config_dir`, apply its defaults list and the overrides."""
    cwd = cwd or os.getcwd()
    name = config_name if config_name.endswith((".yaml", ".yml")) else config_name + ".yaml"
    with open(os.path.join(config_dir, name)) as fh:
        primary = yaml.safe_load(fh) or {}
    defaults = primary.pop("defaults", []) or []
    groups = {}
    for item in defaults:
        if isinstance(item, dict):
            groups.update(item)
    values, deletions = [], []
    for ov in overrides:
        if "=" not in ov:
            raise ValueError(f"override '{ov}' is not of the form key=value")
        key, raw = ov.split("=", 1)
        key = key.lstrip("+")
        if key.startswith("~"):
            deletions.append(key[1:])
            continue
        if key in groups and isinstance(raw, str) and not raw.startswith(("{", "[")):
            groups[key] = raw                    # config-group switch, e.g. data_paths=nusc.yaml
        else:
            values.append((key, yaml.safe_load(raw) if raw != "" else ""))
    cfg = {}
    for group, choice in groups.items():
        if choice in (None, "null"):
            continue
        fname = choice if str(choice).endswith((".yaml", ".yml")) else f"{choice}.yaml"
        with open(os.path.join(config_dir, group, fname)) as fh:
            cfg[group] = yaml.safe_load(fh) or {}
    for k, v in primary.items():
        cfg[k] = v
    for key, v in values:
        _set_dotted(cfg, key, v)
    for key in deletions:
        cfg.pop(key, None)
    hydra_vars = {"runtime.cwd": cwd, "run.dir": run_dir or cwd}
    return _wrap(_resolve_value(cfg, cfg, hydra_vars))


def main(config_path=None, config_name=None, **_ignored):
    """Decorator with hydra.main's calling convention."""

    def deco(fn):
        @functools.wraps(fn)
        def wrapper(cfg_passthrough=None):
            if cfg_passthrough is not None:
                return fn(cfg_passthrough)
            src_dir = os.path.dirname(os.path.abspath(inspect.getsourcefile(fn)))
            cfg_dir = os.path.normpath(os.path.join(src_dir, config_path or "."))
            cwd = os.getcwd()
            overrides = [a for a in sys.argv[1:] if not a.startswith("-")]
            explicit = [a.split("=", 1)[1] for a in overrides if a.startswith("hydra.run.dir=")]
            overrides = [a for a in overrides if not a.startswith("hydra.")]
            now = datetime.datetime.now()
            run_dir = explicit[0] if explicit else os.path.join(
                "outputs", now.strftime("%Y-%m-%d"), now.strftime("%H-%M-%S"))
            cfg = compose(cfg_dir, config_name, overrides, cwd=cwd, run_dir=run_dir)
            os.makedirs(run_dir, exist_ok=True)
            os.chdir(run_dir)                    # Hydra 1.0/1.1 behaviour the reference was written for
            try:
                return fn(cfg)
            finally:
                os.chdir(cwd)
        return wrapper
    return deco


def get_hydra():
    """(main_decorator, DictConfig, OmegaConf) from the real packages when present."""
    try:
        import hydra  # type: ignore
        from omegaconf import DictConfig as DC, OmegaConf as OC  # type: ignore
        return hydra.main, DC, OC
    except Exception:
        return main, DictConfig, OmegaConf
