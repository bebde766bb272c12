"""Config parsing with the reference's conventions (`parse_utils.py:6-77`): every key of every section of the
.ini is flattened into one namedtuple (section names are ignored), values are typed by shape — number, bool,
None, comma list, else string — and `--override "k=v,k2=v2"` strings are parsed by `parse_value`."""
from __future__ import annotations

import configparser
from ast import literal_eval
from collections import namedtuple


def _looks_numeric(value: str) -> bool:
    # decimal / exponent formats, as the reference's test: strip one '.', '+', '-', 'e' and require digits
    return value.replace(".", "", 1).replace("+", "", 1).replace("-", "", 1).replace("e", "", 1).isdigit()


def _parse_list(value: str, sep: str, strip: bool):
    items = value.split(sep)
    numeric = any(ch.isdigit() for ch in items[0])
    if "" in items:
        items.remove("")
    if numeric:
        return [literal_eval(v) for v in items]
    return [v.strip() if strip else v for v in items]


def typed(value: str, ini: bool):
    """Type one raw string.  ini=True: .ini semantics (accepts lowercase true/false, ', ' lists);
    ini=False: --override semantics (`parse_value`)."""
    if _looks_numeric(value):
        return literal_eval(value)
    if value in ("True", "False") or (ini and value in ("true", "false")):
        return value in ("True", "true")
    if value == "None":
        return None
    if "," in value:
        if ini:
            return _parse_list(value, ", " if ", " in value else ",", strip=False)
        return _parse_list(value, ",", strip=True)
    return value


def parse_ini(config_path: str):
    reader = configparser.ConfigParser()
    reader.read(config_path)
    data = {}
    for section in reader.sections():
        for key, value in reader.items(section):
            data[key] = typed(value, ini=True)
    return namedtuple("Config", list(data.keys()))(**data)


def parse_value(value: str):
    return typed(value, ini=False)


def parse_override(text: str) -> dict:
    """`--override` grammar of main.py:24-45: "k=v" or "k=v,k2=v2,..." where a value may itself be a comma list
    (the key of the next pair is whatever follows the last comma before the next '=')."""
    parts = text.split("=")
    assert len(parts) > 1
    if len(parts) == 2:
        return {parts[0]: parse_value(parts[1])}
    keys = [parts[0]] + [p.split(",")[-1] for p in parts[1:-1]]
    values = [p.replace("," + k, "") for p, k in zip(parts[1:-1], keys[1:])] + [parts[-1]]
    values = [v.replace("[", "").replace("]", "") for v in values]
    return {k: parse_value(v) for k, v in zip(keys, values)}


def apply_override(cfg, override: dict):
    merged = cfg._asdict()
    merged.update(override)
    return namedtuple("Config", tuple(merged.keys()))(**merged)


def save_effective_config(cfg, path: str):
    """Effective config as one [Default] section (main.py:47-59): lists are written without brackets."""
    out = configparser.ConfigParser()
    out.add_section("Default")
    for key, value in cfg._asdict().items():
        text = str(value)
        out["Default"][key] = text if key == "name" else text.replace("[", "").replace("]", "")
    with open(path, "w") as f:
        out.write(f)
