"""TEST / BENCH INFRASTRUCTURE.  Imports the unmodified reference (bytecode under ``oracle/_ref/``, see build_ref.py)
as a private set of modules: ``utils``, ``environment``, ``renderers``, ``losses`` exactly as the reference's own scripts
import them by bare name, with the one import that is absent everywhere (``pyredner``, only used by the out-of-scope
RednerRenderer) stubbed.  Nothing in ``svbrdf_estimation_b200`` imports this."""
import importlib.machinery
import importlib.util
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
_cache = None


def available():
    try:
        with open(os.path.join(REF, "PYTHON_VERSION")) as f:
            if f.read().strip() != "%d.%d" % sys.version_info[:2]:
                return False
    except OSError:
        return False
    return all(os.path.exists(os.path.join(REF, m + ".bin")) for m in ("utils", "environment", "renderers", "losses"))


def load():
    """-> namespace with .utils .environment .renderers .losses (the reference's modules), or None."""
    global _cache
    if _cache is not None:
        return _cache
    if not available():
        return None
    names = ("utils", "environment", "renderers", "losses")
    saved = {n: sys.modules.get(n) for n in names + ("pyredner", "cv2")}
    mods = {}
    try:
        if "pyredner" not in sys.modules:
            sys.modules["pyredner"] = types.ModuleType("pyredner")
        try:
            import cv2  # noqa: F401  (renderers.py:1; only the visualisation helper uses it)
        except Exception:
            sys.modules["cv2"] = types.ModuleType("cv2")
        for n in names:
            loader = importlib.machinery.SourcelessFileLoader(n, os.path.join(REF, n + ".bin"))
            spec = importlib.util.spec_from_loader(n, loader)
            m = importlib.util.module_from_spec(spec)
            sys.modules[n] = m                      # the reference's modules import each other by bare name
            loader.exec_module(m)
            mods[n] = m
    finally:
        for n, m in saved.items():
            if m is None:
                sys.modules.pop(n, None)
            else:
                sys.modules[n] = m
    _cache = types.SimpleNamespace(**mods)
    return _cache


def rendering_loss_and_grad(input_maps, target_maps, configs, n_random=None):
    """RenderingLoss(LocalRenderer()) forward + backward of the unmodified reference on given scene records
    [B,N,9] (its two sampler functions are replaced for the call, everything else is the reference's code)."""
    import torch
    ref = load()
    env = ref.environment
    state = {"b": 0}
    if n_random is None:
        n_random = configs.shape[1] // 3            # 3 of 9, 9 of 27: the split RenderingLoss samples (losses.py:26-27)

    def scenes(rows):
        return [env.Scene(env.Camera([float(v) for v in r[0:3]]), env.Light([float(v) for v in r[3:6]], [float(v) for v in r[6:9]])) for r in rows]

    def random_scenes(count):
        return scenes(configs[state["b"], :count])

    def specular_scenes(count):
        rows = configs[state["b"], configs.shape[1] - count:]
        state["b"] += 1
        return scenes(rows)
    keep = (env.generate_random_scenes, env.generate_specular_scenes)
    loss_mod = ref.losses.RenderingLoss(ref.renderers.LocalRenderer())
    loss_mod.random_configuration_count = n_random
    loss_mod.specular_configuration_count = configs.shape[1] - n_random
    x = input_maps.detach().clone().requires_grad_(True)
    try:
        env.generate_random_scenes, env.generate_specular_scenes = random_scenes, specular_scenes
        loss = loss_mod(x, target_maps)
        loss.backward()
    finally:
        env.generate_random_scenes, env.generate_specular_scenes = keep
    return loss.detach(), x.grad
