"""Namespace twin of ``pytorch3d.renderer``: everything the reference scripts import from it
(renderer.py:13-26, torch_renderer.py:14-31, camera_pose_optimizer.py:25-43, mesh_deformer.py:26-40)."""
from .blending import BlendParams  # noqa: F401
from .cameras import (  # noqa: F401
    CamerasBase, FoVOrthographicCameras, FoVPerspectiveCameras, OpenGLOrthographicCameras,
    OpenGLPerspectiveCameras, OrthographicCameras, PerspectiveCameras, SfMOrthographicCameras,
    SfMPerspectiveCameras, camera_position_from_spherical_angles, get_world_to_view_transform,
    look_at_rotation, look_at_view_transform)
from .lighting import AmbientLights, DirectionalLights, Materials, PointLights  # noqa: F401
from .clip import (  # noqa: F401
    ClipFrustum, ClippedFaces, clip_faces, convert_clipped_rasterization_to_original_faces)
from .rasterizer import (  # noqa: F401
    Fragments, MeshRasterizer, RasterizationSettings, rasterize_meshes, set_fragment_cache,
    set_near_plane_clipping)
from .points_renderer import (  # noqa: F401
    AlphaCompositor, NormWeightedCompositor, PointFragments, PointsRasterizationSettings, PointsRasterizer,
    PointsRenderer, alpha_composite, norm_weighted_sum, rasterize_points)
from .shader import (  # noqa: F401
    HardPhongShader, MeshRenderer, MeshRendererWithFragments, SoftPhongShader, SoftSilhouetteShader,
    TexturedSoftPhongShader)
from .textures import TexturesUV, TexturesVertex  # noqa: F401
