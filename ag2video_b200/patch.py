"""Rebind the reference's hot-path names to the sm_100a operators.

The reference binds its operators with ``from ... import`` at module scope
(model.py:4, generator.py:4, discriminator.py:6-9, architecture.py:11,
spade_generator.py:5), so a replacement has to be written into every importing
module (or into ``sys.modules`` before those imports).  ``install()`` does both:
modules already imported are patched in place, and the defining modules are
patched so later ``from ... import`` statements pick the new objects up.
"""
import importlib
import sys

from . import bilinear as _bilinear
from . import graph as _graph
from . import layout as _layout
from . import spade as _spade

# (defining module, name) -> replacement
_TABLE = {
    ('models.graph_models.graph', 'GraphTripleConv'): _graph.GraphTripleConv,
    ('models.layout', 'boxes_to_layout'): _layout.boxes_to_layout,
    ('models.layout', 'masks_to_layout'): _layout.masks_to_layout,
    ('models.bilinear', 'crop_bbox_batch'): _bilinear.crop_bbox_batch,
    ('models.bilinear', 'crop_bbox'): _bilinear.crop_bbox,
    ('models.spade_models.networks.normalization', 'SPADE'): _spade.SPADE,
    ('models.spade_models.networks.architecture', 'SPADEResnetBlock'): _spade.SPADEResnetBlock,
}

# modules that import those names by value
_IMPORTERS = [
    'models.graph_models.model', 'models.spade_models.networks.discriminator',
    'models.spade_models.networks.generator', 'models.spade_models.networks.architecture',
    'models.spade_models.networks.flows_generator', 'models.spade_models.networks.spade_generator',
]


def install(import_missing=True):
    """Returns the list of (module, name) bindings that were replaced."""
    done = []
    originals = {}
    for (mod_name, name), repl in _TABLE.items():
        mod = sys.modules.get(mod_name)
        if mod is None and import_missing:
            try:
                mod = importlib.import_module(mod_name)
            except Exception:
                mod = None
        if mod is None or not hasattr(mod, name):
            continue
        originals[name] = getattr(mod, name)
        setattr(mod, name, repl)
        done.append((mod_name, name))
    for mod_name in _IMPORTERS:
        mod = sys.modules.get(mod_name)
        if mod is None:
            continue
        for name, orig in originals.items():
            if getattr(mod, name, None) is orig:
                setattr(mod, name, _TABLE[next(k for k in _TABLE if k[1] == name)])
                done.append((mod_name, name))
    return done
