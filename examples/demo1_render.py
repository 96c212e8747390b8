"""Rendering a textured mesh from a circle of viewpoints and with a sweep of sigma / gamma: port of the reference's
demo1-render.py (/root/reference/demo1-render.py:12-60) to jrender_b200 -- same loader call, renderer defaults, camera path
and parameter sweep.  Frames are written as PNG files (the reference writes GIFs through imageio, which is not installed
here).

    python examples/demo1_render.py -i data/obj/spot/spot_triangulated.obj -o out/            # needs the .obj/.mtl/.png
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jrender_b200 as jr  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT_OBJ = os.path.join(ROOT, "baseline", "_ref", "assets", "data", "obj", "spot", "spot_triangulated.obj")


def save_png(path, chw):
    import cv2
    img = (255 * chw.detach().clamp(0, 1).cpu().numpy().transpose(1, 2, 0)).astype(np.uint8)
    cv2.imwrite(path, img[:, :, ::-1])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('-i', '--filename-input', type=str, default=DEFAULT_OBJ)
    ap.add_argument('-o', '--output-dir', type=str, default=os.path.join(ROOT, 'gpurun_out', 'output_render'))
    ap.add_argument('--step', type=int, default=4, help="azimuth step in degrees (demo1: 4)")
    args = ap.parse_args()
    camera_distance, elevation = 2.732, 30
    dev = torch.device('cuda:0')

    # load from Wavefront .obj file (demo1-render.py:27); create renderer with SoftRas (:29)
    mesh = jr.Mesh.from_obj(args.filename_input, load_texture=True, texture_res=5, texture_type='surface', dr_type='softras').to(dev)
    renderer = jr.Renderer(dr_type='softras')
    os.makedirs(args.output_dir, exist_ok=True)

    # draw object from different views (:33-45)
    for num, azimuth in enumerate(range(0, 360, args.step)):
        mesh.reset_()
        renderer.transform.set_eyes_from_angles(camera_distance, elevation, azimuth)
        rgb = renderer.render_mesh(mesh, mode='rgb')
        save_png(os.path.join(args.output_dir, 'rotation_%03d.png' % num), rgb[0])

    # draw object with different sigma and gamma (:47-60)
    renderer.transform.set_eyes_from_angles(camera_distance, elevation, 45)
    for num, gamma_pow in enumerate(np.arange(-4, -2, 0.2)):
        mesh.reset_()
        renderer.set_gamma(10 ** gamma_pow)
        renderer.set_sigma(10 ** (gamma_pow - 1))
        images = renderer.render_mesh(mesh, mode='rgb')
        save_png(os.path.join(args.output_dir, 'bluring_%03d.png' % num), images[0])
    print("wrote %d frames to %s" % (len(os.listdir(args.output_dir)), args.output_dir))


if __name__ == '__main__':
    main()
