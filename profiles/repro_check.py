import sys, torch
sys.path[:0] = ['.', 'tests']
import torch_renderer_b200 as trb
from helpers import uv_sphere
DEV = torch.device("cuda:0")
trb.set_fragment_cache(False)
v, f = uv_sphere(40, 60, 1.0, noise=0.04, seed=5)
R, T = trb.look_at_view_transform(1.9, torch.tensor([15.0, -35.0]), torch.tensor([30.0, 190.0]))
cols = torch.rand(1, v.shape[0], 3, device=DEV)
for lights in ("point", "ambient"):
    mesh = trb.Meshes([v.to(DEV)], [f.to(DEV)], textures=trb.TexturesVertex(cols)).extend(2)
    cams = trb.FoVPerspectiveCameras(device=DEV, R=R.to(DEV), T=T.to(DEV))
    lt = trb.PointLights(device=DEV, location=[[0.0, 2.0, -3.0]]) if lights == "point" else trb.AmbientLights(device=DEV)
    rend = trb.MeshRendererWithFragments(
        trb.MeshRasterizer(cams, trb.RasterizationSettings(image_size=(208, 256), blur_radius=0.0, faces_per_pixel=1)),
        trb.SoftPhongShader(device=DEV, cameras=cams, lights=lt))
    img0, fr0 = rend(mesh)
    mx = 0.0; nd = 0
    for _ in range(30):
        img, fr = rend(mesh)
        d = (img - img0).abs()
        mx = max(mx, float(d.max())); nd = max(nd, int((d > 0).sum()))
        assert torch.equal(fr.pix_to_face, fr0.pix_to_face)
    print(lights, "max abs image diff", mx, "pixels differing", nd)
