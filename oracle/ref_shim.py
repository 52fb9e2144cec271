"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference hot path.

Imports `pipeline/causal_inference.py`, `utils/wan_wrapper.py`, `utils/scheduler.py` and
`wan/modules/{causal_model,model,attention}.py` straight from the read-only reference
checkout (default `/root/reference`) so that the CPU restatement in `oracle/` can be pinned
against the real thing and golden vectors can be generated (`oracle/make_golden.py`).

Nothing here edits or copies reference source.  The reference needs a handful of packages
that are not in this image (diffusers, ftfy, easydict ...) and evaluates
`torch.cuda.current_device()` at import time in two modules, so we register small stand-ins in
`sys.modules` *before* importing it (SURVEY.md section 8c lists them):

  1. `diffusers.configuration_utils.{ConfigMixin,register_to_config}` and
     `diffusers.models.modeling_utils.ModelMixin`  (base classes only; causal_model.py:13-15)
  2. empty package shells for `wan`, `wan.modules`, `pipeline`, `utils`, `demo_utils` with
     `__path__` pointing into the checkout (bypasses their heavy `__init__`s)
  3. stubs for `ftfy`, `wan.modules.t5`, `wan.modules.vae`, `wan.modules.tokenizers`,
     `demo_utils.memory`, `utils.lora`
  3b. for the 50-step sampler (`pipeline/causal_diffusion_inference.py`, `wan/utils/fm_solvers_unipc.py`):
     `diffusers.schedulers.scheduling_utils.{SchedulerMixin,SchedulerOutput,KarrasDiffusionSchedulers}`,
     `diffusers.utils.{deprecate,is_scipy_available}`, `diffusers.utils.torch_utils.randn_tensor`, a
     `register_to_config` that records the constructor arguments in `self.config` (the solver reads
     `self.config.solver_order` etc.), and a `wan.modules.clip` stub (needs torchvision + xlm_roberta)
  4. CPU only: cross-attention calls `flash_attention` directly (model.py:189), which asserts
     CUDA before reaching its own SDPA fallback (attention.py:62 vs :68) -> rebind to
     `attention.attention` with the FA flags cleared.

The reference does not exist on the GPU box; only this container can call `load_reference()`.
Product code never imports this module.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import torch
from torch import nn

REF_ROOT = os.environ.get("SFB_REFERENCE_ROOT", "/root/reference")


def reference_available(root: str = REF_ROOT) -> bool:
    return os.path.isfile(os.path.join(root, "pipeline", "causal_inference.py"))


def _shell(name: str, path: str | None) -> types.ModuleType:
    mod = types.ModuleType(name)
    if path is not None:
        mod.__path__ = [path]
    sys.modules[name] = mod
    return mod


def _install_stubs(root: str) -> None:
    # (1) diffusers base classes
    diffusers = _shell("diffusers", None)
    diffusers.__path__ = []
    cu = _shell("diffusers.configuration_utils", None)

    class ConfigMixin:  # noqa: D401 - stand-in
        pass

    def register_to_config(fn):
        """Like diffusers': bind the constructor arguments (defaults included) to `self.config`."""
        import functools
        import inspect
        sig = inspect.signature(fn)

        @functools.wraps(fn)
        def init(self, *args, **kwargs):
            bound = sig.bind(self, *args, **kwargs)
            bound.apply_defaults()
            cfg = {k: v for k, v in bound.arguments.items() if k != "self"}
            fn(self, *args, **kwargs)
            object.__setattr__(self, "config", types.SimpleNamespace(**cfg))
        return init

    cu.ConfigMixin = ConfigMixin
    cu.register_to_config = register_to_config
    models = _shell("diffusers.models", None)
    models.__path__ = []
    mu = _shell("diffusers.models.modeling_utils", None)

    class ModelMixin(nn.Module):
        pass

    mu.ModelMixin = ModelMixin
    diffusers.configuration_utils = cu
    diffusers.models = models
    models.modeling_utils = mu
    # (3b) scheduler base classes of the UniPC / DPM++ flow solvers
    sched = _shell("diffusers.schedulers", None)
    sched.__path__ = []
    su = _shell("diffusers.schedulers.scheduling_utils", None)

    class SchedulerMixin:
        pass

    class SchedulerOutput:
        def __init__(self, prev_sample):
            self.prev_sample = prev_sample

    su.SchedulerMixin, su.SchedulerOutput, su.KarrasDiffusionSchedulers = SchedulerMixin, SchedulerOutput, []
    du = _shell("diffusers.utils", None)
    du.__path__ = []
    du.deprecate = lambda *a, **k: None
    du.is_scipy_available = lambda: False
    tu = _shell("diffusers.utils.torch_utils", None)
    tu.randn_tensor = lambda shape, generator=None, device=None, dtype=None: torch.randn(
        shape, generator=generator, device=device, dtype=dtype)
    diffusers.schedulers, sched.scheduling_utils, diffusers.utils, du.torch_utils = sched, su, du, tu

    # (2) package shells
    _shell("wan", os.path.join(root, "wan"))
    _shell("wan.modules", os.path.join(root, "wan", "modules"))
    _shell("wan.utils", os.path.join(root, "wan", "utils"))
    _shell("pipeline", os.path.join(root, "pipeline"))
    _shell("utils", os.path.join(root, "utils"))
    _shell("demo_utils", os.path.join(root, "demo_utils"))

    # (3) stubs for modules that need absent deps / a CUDA device at import time
    _shell("ftfy", None)
    t5 = _shell("wan.modules.t5", None)
    t5.umt5_xxl = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("t5 stub"))
    vae = _shell("wan.modules.vae", None)
    vae._video_vae = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("vae stub"))
    clip = _shell("wan.modules.clip", None)
    clip.CLIPModel = object
    tok = _shell("wan.modules.tokenizers", None)
    tok.HuggingfaceTokenizer = object
    mem = _shell("demo_utils.memory", None)
    mem.gpu = torch.device("cpu")
    mem.get_cuda_free_memory_gb = lambda *a, **k: 0.0
    mem.DynamicSwapInstaller = object
    mem.move_model_to_device_with_memory_preservation = lambda *a, **k: None
    lora = _shell("utils.lora", None)
    lora.apply_lora = lambda *a, **k: 0
    lora.load_lora_weights = lambda *a, **k: (0, 0)


_LOADED = None


def load_reference(root: str = REF_ROOT, force_sdpa: bool | None = None):
    """Return a namespace with the reference classes of the hot path.

    force_sdpa: route attention through the reference's own SDPA branch (attention.py:187-202).
    Defaults to True when no CUDA device is present.
    """
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not reference_available(root):
        raise FileNotFoundError(f"reference checkout not found at {root}")
    if force_sdpa is None:
        force_sdpa = not torch.cuda.is_available()
    _install_stubs(root)
    attention = importlib.import_module("wan.modules.attention")
    if force_sdpa:
        attention.FLASH_ATTN_2_AVAILABLE = False
        attention.FLASH_ATTN_3_AVAILABLE = False
    model = importlib.import_module("wan.modules.model")
    if force_sdpa:
        model.flash_attention = attention.attention
    causal_model = importlib.import_module("wan.modules.causal_model")
    scheduler = importlib.import_module("utils.scheduler")
    wan_wrapper = importlib.import_module("utils.wan_wrapper")
    causal_inference = importlib.import_module("pipeline.causal_inference")
    unipc = importlib.import_module("wan.utils.fm_solvers_unipc")
    causal_diffusion = importlib.import_module("pipeline.causal_diffusion_inference")

    ns = types.SimpleNamespace(
        attention=attention, model=model, causal_model=causal_model, scheduler=scheduler,
        wan_wrapper=wan_wrapper, causal_inference=causal_inference,
        CausalWanModel=causal_model.CausalWanModel,
        WanModel=model.WanModel,
        WanDiffusionWrapper=wan_wrapper.WanDiffusionWrapper,
        FlowMatchScheduler=scheduler.FlowMatchScheduler,
        CausalInferencePipeline=causal_inference.CausalInferencePipeline,
        FlowUniPCMultistepScheduler=unipc.FlowUniPCMultistepScheduler,
        CausalDiffusionInferencePipeline=causal_diffusion.CausalDiffusionInferencePipeline,
    )
    _LOADED = ns
    return ns


def load_reference_vae(root: str = REF_ROOT):
    """The reference's real `wan/modules/vae.py` (torch + einops only), loaded under a private module name because
    `wan.modules.vae` itself is stubbed above for the wrapper import."""
    import importlib.util
    name = "_sfb_reference_vae"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(root, "wan", "modules", "vae.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference_vae_block3(root: str = REF_ROOT):
    """`demo_utils/vae_block3.py` (the streaming decoder with an explicit feature-cache list); it imports its layers from
    `wan.modules.vae`, so the real module stands in for the stub while it loads."""
    import importlib.util
    name = "_sfb_reference_vae_block3"
    if name in sys.modules:
        return sys.modules[name]
    real = load_reference_vae(root)
    if "wan" not in sys.modules:
        _shell("wan", os.path.join(root, "wan"))
        _shell("wan.modules", os.path.join(root, "wan", "modules"))
    stub = sys.modules.get("wan.modules.vae")
    sys.modules["wan.modules.vae"] = real
    try:
        spec = importlib.util.spec_from_file_location(name, os.path.join(root, "demo_utils", "vae_block3.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
    finally:
        if stub is not None:
            sys.modules["wan.modules.vae"] = stub
        else:
            sys.modules.pop("wan.modules.vae", None)
    return mod


def load_reference_t5(root: str = REF_ROOT):
    """The reference's real `wan/modules/t5.py`, loaded under a private name inside the `wan.modules` package shell (it
    imports `.tokenizers`, stubbed above).  Its `T5EncoderModel` class evaluates `torch.cuda.current_device()` as a
    default argument at import time, so that call is answered with 0 while the module loads."""
    import importlib.util
    name = "wan.modules._sfb_reference_t5"
    if name in sys.modules:
        return sys.modules[name]
    if "wan.modules.tokenizers" not in sys.modules:
        load_reference(root)
    spec = importlib.util.spec_from_file_location(name, os.path.join(root, "wan", "modules", "t5.py"))
    mod = importlib.util.module_from_spec(spec)
    mod.__package__ = "wan.modules"
    sys.modules[name] = mod
    real = torch.cuda.current_device
    torch.cuda.current_device = lambda: 0
    try:
        spec.loader.exec_module(mod)
    finally:
        torch.cuda.current_device = real
    return mod


def make_reference_wrapper(ref, model_cfg: dict, timestep_shift: float, seed: int = 0,
                           dtype=torch.bfloat16):
    """Random-init `WanDiffusionWrapper` without `from_pretrained` (wan_wrapper.py:139-147).

    The reference zero-inits `head.head.weight` and every Linear bias and sets the RMSNorm
    affines to one (causal_model.py:1113-1128, model.py:76), which would make `flow_pred == 0`
    and parity vacuous, so those parameters are re-randomised.
    """
    WanDiffusionWrapper = ref.WanDiffusionWrapper

    class RandomInitWrapper(WanDiffusionWrapper):
        def __init__(self):
            nn.Module.__init__(self)
            torch.manual_seed(seed)
            self.model = ref.CausalWanModel(**model_cfg)
            rerandomise_(self.model, seed + 1)
            self.model = self.model.to(dtype).eval()
            self.uniform_timestep = False
            self.scheduler = ref.FlowMatchScheduler(
                shift=timestep_shift, sigma_min=0.0, extra_one_step=True)
            self.scheduler.set_timesteps(1000, training=True)
            self.seq_len = 32760
            self.post_init()

    return RandomInitWrapper()


@torch.no_grad()
def rerandomise_(model: nn.Module, seed: int) -> None:
    """In-place: N(0, .02) biases / head weight, 1 + N(0, .02) norm affines."""
    g = torch.Generator().manual_seed(seed)
    for name, p in model.named_parameters():
        if name.endswith(".bias"):
            p.copy_(torch.randn(p.shape, generator=g) * 0.02)
        elif name == "head.head.weight":
            p.copy_(torch.randn(p.shape, generator=g) * 0.02)
        elif ("norm_q" in name or "norm_k" in name or "norm3" in name) and name.endswith(".weight"):
            p.copy_(1.0 + torch.randn(p.shape, generator=g) * 0.02)
