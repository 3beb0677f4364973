"""Import-time stand-in for matplotlib (simulator.py:8-10, common/visualization.py:3-7 import it at module level).
With `"render": false` in the simulation config nothing is drawn; any drawing call on the stub raises."""
import types


class _Unavailable:
    def __init__(self, name):
        self._name = name

    def __call__(self, *a, **k):
        raise RuntimeError("matplotlib is not installed: %s is unavailable (run with \"render\": false)" % self._name)

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        return _Unavailable(self._name + "." + item)


def modules():
    mpl = types.ModuleType("matplotlib"); mpl.__path__ = []
    mpl.use = lambda *a, **k: None
    plt = types.ModuleType("matplotlib.pyplot")

    def plt_attr(name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Unavailable("matplotlib.pyplot." + name)
    plt.__getattr__ = plt_attr
    patches = types.ModuleType("matplotlib.patches")
    patches.Circle, patches.Ellipse = _Unavailable("Circle"), _Unavailable("Ellipse")
    tk = types.ModuleType("mpl_toolkits"); tk.__path__ = []
    m3 = types.ModuleType("mpl_toolkits.mplot3d"); m3.__path__ = []
    art = types.ModuleType("mpl_toolkits.mplot3d.art3d")
    art.Poly3DCollection = _Unavailable("Poly3DCollection")
    mpl.pyplot, mpl.patches, tk.mplot3d, m3.art3d = plt, patches, m3, art
    return {"matplotlib": mpl, "matplotlib.pyplot": plt, "matplotlib.patches": patches, "mpl_toolkits": tk,
            "mpl_toolkits.mplot3d": m3, "mpl_toolkits.mplot3d.art3d": art}
