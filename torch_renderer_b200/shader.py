"""``SoftPhongShader``, ``HardPhongShader``, ``SoftSilhouetteShader`` and ``MeshRenderer`` with the
PyTorch3D call surface.  Reference usage: renderer.py:87-101,
torch_renderer.py:102-108,144-158, camera_pose_optimizer.py:130-158, mesh_deformer.py:142-145,197,
myrenderer.py:88,105, batch_rendering_test.py:171-200.

Each shader call is ONE fused CUDA kernel (csrc/shade.cu): attribute interpolation, Phong
lighting and blending; the backward is one kernel as well.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch
import torch.nn as nn

from . import _lib, ops
from .blending import BlendParams
from .common import named_tensors
from .lighting import AmbientLights, DirectionalLights, Materials, PointLights
from .rasterizer import Fragments
from .structures import Meshes
from .textures import TexturesVertex

_LIGHT_KIND = {"ambient": _lib.LIGHT_AMBIENT, "point": _lib.LIGHT_POINT, "directional": _lib.LIGHT_DIRECTIONAL}


def _rows(t: torch.Tensor, N: int, device, what: str) -> torch.Tensor:
    """(n, C) with n in {1, N} -> (N, C) on device."""
    if t.device != device:
        t = t.to(device)
    if t.dim() == 1:
        t = t[:, None]
    if t.shape[0] == N:
        return t
    if t.shape[0] == 1:
        return t.expand(N, t.shape[1])
    raise ValueError(f"{what} has batch size {t.shape[0]}; expected 1 or {N}")


def _scalar_rows(v, N: int, device) -> torch.Tensor:
    if not torch.is_tensor(v):
        return torch.full((N, 1), float(v), dtype=torch.float32, device=device)
    return _rows(v.reshape(-1, 1).float(), N, device, "znear/zfar")


def _colour_products(N, device, lights, materials):
    """(ambient, diffuse, specular) = material colour * light colour, f32 (N, 3) each (zeros for the terms an
    ambient-only light does not have).  Differentiable torch ops: the kernels see the products only."""
    amb = _rows(materials.ambient_color, N, device, "materials") * _rows(lights.ambient_color, N, device, "lights")
    if getattr(lights, "kind", None) == "ambient":
        zeros3 = torch.zeros((N, 3), dtype=torch.float32, device=device)
        return amb, zeros3, zeros3
    dif = _rows(materials.diffuse_color, N, device, "materials") * _rows(lights.diffuse_color, N, device, "lights")
    spec = _rows(materials.specular_color, N, device, "materials") * _rows(lights.specular_color, N, device, "lights")
    return amb, dif, spec


class _ColourGradFn(torch.autograd.Function):
    """Gradients w.r.t. the light / material COLOURS, without touching the kernels: the blended RGB is linear in
    the three colour products (colour_k = ambient * texel + diffuse * texel * cos + specular * a^shininess, and the
    blend weights do not depend on colour), channel by channel.  So d rgb / d ambient is the image shaded with
    ambient = 1, diffuse = specular = background = 0, and likewise for the other two: three extra shade-only
    passes over the stored Fragments in the backward, run only when a colour requires a gradient.  Forward is the
    identity on ``images`` (the kernels' own parameter-block gradient is zero in the colour slots)."""

    @staticmethod
    def forward(ctx, images, amb, dif, spec, shade_args):
        ctx.shade_args = shade_args
        return images

    @staticmethod
    def backward(ctx, g_images):
        (bary, zbuf, dists, verts, normals, colors, texels, vp, p2f, faces, table, cfg) = ctx.shade_args
        need = ctx.needs_input_grad
        g_rgb = g_images[..., :3]
        unit = _lib.ShadeConfig()
        ctypes.memmove(ctypes.byref(unit), ctypes.byref(cfg), ctypes.sizeof(cfg))
        unit.background[0] = unit.background[1] = unit.background[2] = 0.0
        grads = []
        for i, lo in enumerate((3, 6, 9)):           # slots of ambient / diffuse / specular in the parameter block
            if not need[1 + i] or (lo > 3 and cfg.light_kind == _lib.LIGHT_AMBIENT):
                grads.append(None)
                continue
            vp_i = vp.detach().clone()
            vp_i[:, 3:12] = 0.0
            vp_i[:, lo:lo + 3] = 1.0
            img = ops.shade(bary.detach(), zbuf.detach(), dists.detach(), None if verts is None else verts.detach(),
                            None if normals is None else normals.detach(),
                            None if colors is None else colors.detach(),
                            None if texels is None else texels.detach(), vp_i, p2f, faces, table, unit)
            grads.append((g_rgb * img[..., :3]).sum(dim=(1, 2)))
        return (g_images, *grads, None)


def _colours_require_grad(lights, materials) -> bool:
    names = ("ambient_color",) if lights.kind == "ambient" else ("ambient_color", "diffuse_color", "specular_color")
    return torch.is_grad_enabled() and any(getattr(o, n).requires_grad for o in (lights, materials) for n in names)


def _with_colour_grads(images, lights, materials, shade_args):
    """Hooks ``_ColourGradFn`` behind ``images`` when a light / material colour requires a gradient."""
    if not _colours_require_grad(lights, materials):
        return images
    N = images.shape[0]
    amb, dif, spec = _colour_products(N, images.device, lights, materials)
    return _ColourGradFn.apply(images, amb, dif, spec, shade_args)


def _view_params(N, device, lights, materials, cameras, znear, zfar, with_camera_center=True) -> torch.Tensor:
    """f32 [N, 20] parameter block of the shade kernel (layout: include/trb.h).  With
    ``with_camera_center=False`` the camera-centre slots are left zero (the fused kernel fills them)."""
    kind = getattr(lights, "kind", None)
    if kind not in _LIGHT_KIND:
        raise ValueError(f"unsupported lights object {type(lights).__name__}")
    zeros3 = torch.zeros((N, 3), dtype=torch.float32, device=device)
    amb, dif, spec = _colour_products(N, device, lights, materials)
    if kind == "ambient":
        vec = zeros3
    else:
        vec = _rows(lights.location if kind == "point" else lights.direction, N, device, "lights")
    if materials.shininess.requires_grad:
        raise NotImplementedError("gradients w.r.t. Materials.shininess are not built (colours, light location / "
                                  "direction and the camera are differentiable)")
    shin = _rows(materials.shininess.reshape(-1, 1), N, device, "materials")
    if kind != "ambient" and with_camera_center:
        cam = _rows(cameras.get_camera_center(), N, device, "cameras")
    else:
        cam = zeros3
    pad = torch.zeros((N, 2), dtype=torch.float32, device=device)
    return torch.cat([vec, amb, dif, spec, shin, cam, _scalar_rows(znear, N, device),
                      _scalar_rows(zfar, N, device), pad], dim=1)


def _cached_view_params(owner, N, device, lights, materials, cameras, znear, zfar, with_camera_center):
    """Memoises the parameter block on the shader while every input is an unmodified constant."""
    # plain attributes, Parameters and buffers alike (``lights.location = nn.Parameter(...)`` lands in _parameters)
    tensors = [v for obj in (lights, materials) for v in named_tensors(obj).values()]
    tensors += [v for v in (znear, zfar) if torch.is_tensor(v)]
    if with_camera_center:
        tensors += [cameras.R, cameras.T]
    if any(t.requires_grad for t in tensors):
        return _view_params(N, device, lights, materials, cameras, znear, zfar, with_camera_center)
    key = (N, str(device), with_camera_center, id(lights), id(materials),
           tuple((id(t), t._version) for t in tensors),
           tuple(float(v) for v in (znear, zfar) if not torch.is_tensor(v)))
    cache = owner.__dict__.get("_trb_vp_cache")
    if cache is None or cache[0] != key:
        cache = (key, _view_params(N, device, lights, materials, cameras, znear, zfar, with_camera_center))
        owner.__dict__["_trb_vp_cache"] = cache
    return cache[1]


def _shade_config(fragments: Fragments, shader: int, light_kind: int, texture_mode: int,
                  blend_params: BlendParams) -> _lib.ShadeConfig:
    N, H, W, K = fragments.pix_to_face.shape
    bg = blend_params.background_color
    bg = [float(x) for x in (bg.tolist() if torch.is_tensor(bg) else bg)]
    cfg = _lib.ShadeConfig()
    cfg.N, cfg.H, cfg.W, cfg.K = N, H, W, K
    cfg.shader, cfg.light_kind, cfg.texture_mode = shader, light_kind, texture_mode
    cfg.sigma, cfg.gamma = float(blend_params.sigma), float(blend_params.gamma)
    cfg.background[0], cfg.background[1], cfg.background[2] = bg[0], bg[1], bg[2]
    return cfg


class _ShaderBase(nn.Module):
    def __init__(self, device="cpu", cameras=None, lights=None, materials=None,
                 blend_params: Optional[BlendParams] = None) -> None:
        super().__init__()
        self.lights = lights if lights is not None else PointLights(device=device)
        self.materials = materials if materials is not None else Materials(device=device)
        self.cameras = cameras
        self.blend_params = blend_params if blend_params is not None else BlendParams()

    def _get_cameras(self, **kwargs):
        cameras = kwargs.get("cameras", self.cameras)
        if cameras is None:
            raise ValueError("Cameras must be specified either at initialization or in the forward pass "
                             "of the shader.")
        return cameras

    def to(self, device):
        if self.cameras is not None:
            self.cameras = self.cameras.to(device)
        self.materials = self.materials.to(device)
        self.lights = self.lights.to(device)
        return self

    def _phong(self, shader: int, fragments: Fragments, meshes: Meshes, **kwargs) -> torch.Tensor:
        cameras = self._get_cameras(**kwargs)
        lights = kwargs.get("lights", self.lights)
        materials = kwargs.get("materials", self.materials)
        blend_params = kwargs.get("blend_params", self.blend_params)
        dev = fragments.pix_to_face.device
        N = fragments.pix_to_face.shape[0]
        table = meshes.view_table()
        if table.N != N:
            raise ValueError("fragments and meshes have different batch sizes")
        textures = meshes.textures
        if textures is None:
            raise ValueError("Meshes does not have textures")
        colors = texels = None
        if isinstance(textures, TexturesVertex):
            shared = table.shared_mesh and textures._replicas == table.N
            if table.shared_mesh and not shared:
                raise ValueError("textures batch does not match the extended mesh")
            colors = textures._unique_features(shared)
            if colors.shape[-1] != 3:
                raise ValueError("Phong shading needs RGB (C=3) vertex features")
            tex_mode = _lib.TEX_VERTEX
        else:
            texels = meshes.sample_textures(fragments)
            tex_mode = _lib.TEX_TEXELS
        znear = kwargs.get("znear", getattr(cameras, "znear", 1.0))
        zfar = kwargs.get("zfar", getattr(cameras, "zfar", 100.0))
        vp = _cached_view_params(self, N, dev, lights, materials, cameras, znear, zfar, True)
        cfg = _shade_config(fragments, shader, _LIGHT_KIND[lights.kind], tex_mode, blend_params)
        args = (fragments.bary_coords, fragments.zbuf, fragments.dists, meshes._unique_verts(),
                meshes._unique_verts_normals(), colors, texels, vp, fragments.pix_to_face,
                meshes.faces_packed_i32(), table, cfg)
        return _with_colour_grads(ops.shade(*args), lights, materials, args)


class SoftPhongShader(_ShaderBase):
    """Per-pixel Phong lighting on interpolated coordinates/normals + softmax_rgb_blend."""

    def forward(self, fragments: Fragments, meshes: Meshes, **kwargs) -> torch.Tensor:
        return self._phong(_lib.SHADER_SOFT_PHONG, fragments, meshes, **kwargs)


class HardPhongShader(_ShaderBase):
    """Per-pixel Phong lighting, closest face only (hard_rgb_blend)."""

    def forward(self, fragments: Fragments, meshes: Meshes, **kwargs) -> torch.Tensor:
        return self._phong(_lib.SHADER_HARD_PHONG, fragments, meshes, **kwargs)


TexturedSoftPhongShader = SoftPhongShader  # legacy PyTorch3D alias


class SoftSilhouetteShader(nn.Module):
    """RGB = 1, alpha = 1 - prod_k (1 - sigmoid(-dist_k / sigma)) (sigmoid_alpha_blend)."""

    def __init__(self, blend_params: Optional[BlendParams] = None) -> None:
        super().__init__()
        self.blend_params = blend_params if blend_params is not None else BlendParams()

    def forward(self, fragments: Fragments, meshes: Meshes, **kwargs) -> torch.Tensor:
        blend_params = kwargs.get("blend_params", self.blend_params)
        cfg = _shade_config(fragments, _lib.SHADER_SOFT_SILHOUETTE, 0, 0, blend_params)
        return ops.shade(fragments.bary_coords, fragments.zbuf, fragments.dists, None, None, None, None,
                         None, fragments.pix_to_face, None, None, cfg)


class MeshRenderer(nn.Module):
    """``images = shader(rasterizer(meshes_world, **kwargs), meshes_world, **kwargs)``.

    When the rasteriser is a ``MeshRasterizer`` and the shader one of this package's shaders on
    per-vertex colours, the whole call is ONE fused C-ABI call (``trb_render_forward``) and its
    backward another; otherwise the two modules are composed as written above."""

    def __init__(self, rasterizer, shader) -> None:
        super().__init__()
        self.rasterizer = rasterizer
        self.shader = shader

    def to(self, device):
        self.rasterizer.to(device)
        self.shader.to(device)
        return self

    def _render_fused(self, meshes_world: Meshes, kwargs, want_fragments: bool = True):
        """Returns (images, Fragments) through the fused pipeline, or None when it does not apply.  With
        ``want_fragments=False`` (``MeshRenderer``: upstream's ``renderer(meshes)`` returns the image and nothing
        else) the kernels write the Fragments of COVERED pixels only -- what the fused backward reads -- instead
        of 28*K bytes for every pixel of the image, and the second element is None."""
        from .rasterizer import MeshRasterizer
        rast, shader = self.rasterizer, self.shader
        if type(rast) is not MeshRasterizer:
            return None
        if type(shader) is SoftSilhouetteShader:
            kind = _lib.SHADER_SOFT_SILHOUETTE
        elif type(shader) is SoftPhongShader:
            kind = _lib.SHADER_SOFT_PHONG
        elif type(shader) is HardPhongShader:
            kind = _lib.SHADER_HARD_PHONG
        else:
            return None
        table = meshes_world.view_table()
        N, dev = table.N, meshes_world.device
        colors = vp = None
        light_kind = 0
        blend_params = kwargs.get("blend_params", shader.blend_params)
        from_rt = False
        phong = kind != _lib.SHADER_SOFT_SILHOUETTE
        tex_map = uv = None
        if phong:
            textures = meshes_world.textures
            if isinstance(textures, TexturesVertex):
                shared = table.shared_mesh and textures._replicas == table.N
                if table.shared_mesh and not shared:
                    return None
                colors = textures._unique_features(shared)
                if colors.shape[-1] != 3 or colors.shape[0] != meshes_world._unique_verts().shape[0]:
                    return None
            else:
                # TexturesUV sampled inside the kernels: one mesh (or one mesh extended to N views) with ONE map
                fused_uv = getattr(textures, "_fused_inputs", None)
                got = None if fused_uv is None else fused_uv(table, meshes_world.faces_packed_i32())
                if got is None:
                    return None
                tex_map, uv = got
        cameras, R, T, proj, spec = rast._resolve(meshes_world, kwargs)   # also stores R, T on `cameras`
        if phong:
            shader_cameras = shader._get_cameras(**kwargs)
            lights = kwargs.get("lights", shader.lights)
            materials = kwargs.get("materials", shader.materials)
            light_kind = _LIGHT_KIND.get(getattr(lights, "kind", None))
            if light_kind is None:
                return None
            if _colours_require_grad(lights, materials):
                return None     # colour gradients hang off the stand-alone shader call (_ColourGradFn)
            # the shader asks its camera object for the centre; when that is the object the rasteriser
            # just used, the centre belongs to this call's R, T and the kernel derives it (and its
            # gradient) itself
            from_rt = shader_cameras is cameras
            znear = kwargs.get("znear", getattr(shader_cameras, "znear", 1.0))
            zfar = kwargs.get("zfar", getattr(shader_cameras, "zfar", 100.0))
            vp = _cached_view_params(shader, N, dev, lights, materials, shader_cameras, znear, zfar, not from_rt)
        # the same scene was just rasterised with the same settings (the reference does that 2-3 times per
        # step): shade the stored Fragments instead of rasterising again
        from .rasterizer import _fragment_cache
        key, tensors = rast._cache_key(meshes_world, spec)
        cached = _fragment_cache.lookup(key, tensors)
        if cached is not None:
            return shader(cached, meshes_world, **kwargs), cached
        # near plane (rasterizer.set_near_plane_clipping): the question is enqueued now and read after the fused
        # render has been enqueued; on "yes" the faces are cut by clip_faces and drawn by the stand-alone rasteriser
        behind = rast._near_plane_question(meshes_world, R, T, spec)
        if behind is not None and spec["cull_to_frustum"]:
            clipped = rast._clipped_fragments(meshes_world, R, T, proj, spec)
            if clipped is not None:
                return shader(clipped, meshes_world, **kwargs), clipped
            behind = None
        bg = blend_params.background_color
        bg = tuple(float(x) for x in (bg.tolist() if torch.is_tensor(bg) else bg))
        token = {"consumed": False}
        if uv is not None:
            spec["uv"] = uv
        # dense Fragments when somebody will look at them: the caller, or the (opt-in) Fragments cache
        sparse = not want_fragments and not _fragment_cache.enabled
        spec.update(_token=token, shader=kind, light_kind=light_kind, sigma=float(blend_params.sigma),
                    gamma=float(blend_params.gamma), background=bg, camera_center_from_rt=from_rt, sparse=sparse)
        images, p2f, zbuf, bary, dists = ops.render(
            meshes_world._unique_verts(), colors, R, T, proj, vp, meshes_world.faces_packed_i32(), table, spec,
            tex_map=tex_map)
        if behind is not None and behind():
            clipped = rast._clipped_fragments(meshes_world, R, T, proj, spec)
            if clipped is not None:
                return shader(clipped, meshes_world, **kwargs), clipped
        if sparse:
            return images, None
        fragments = Fragments(pix_to_face=p2f, zbuf=zbuf, bary_coords=bary, dists=dists)
        _fragment_cache.store(key, tensors, fragments, token)
        return images, fragments

    def forward(self, meshes_world: Meshes, **kwargs) -> torch.Tensor:
        fused = self._render_fused(meshes_world, kwargs, want_fragments=False)
        if fused is not None:
            return fused[0]
        fragments = self.rasterizer(meshes_world, **kwargs)
        return self.shader(fragments, meshes_world, **kwargs)


class MeshRendererWithFragments(MeshRenderer):
    def forward(self, meshes_world: Meshes, **kwargs):
        fused = self._render_fused(meshes_world, kwargs)
        if fused is not None:
            return fused
        fragments = self.rasterizer(meshes_world, **kwargs)
        return self.shader(fragments, meshes_world, **kwargs), fragments
