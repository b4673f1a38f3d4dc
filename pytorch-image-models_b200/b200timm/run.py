"""Launcher for timm scripts that have no import hook (benchmark.py, validate.py, inference.py):

    python -m b200timm.run /path/to/timm/benchmark.py --model vit_base_patch16_224 --bench inference \
        -b 256 --precision bfloat16

installs the plugin (b200timm.plugin.install) and then runs the script unchanged under ``__main__``.
"""
import runpy
import sys


def main():
    if len(sys.argv) < 2:
        raise SystemExit("usage: python -m b200timm.run <timm script.py> [script args...]")
    from . import plugin
    plugin.install()
    script = sys.argv[1]
    sys.argv = sys.argv[1:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
