"""Shim with the reference's module name: `import localize; localize.localize_stanford(cfg, writer, log_dir)`."""
from piccolo_b200.localize import get_init_dict, localize_omniscenes, localize_stanford  # noqa: F401
