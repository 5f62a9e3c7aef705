"""Mirror of the output stage of the reference, `test_utils/export_res.py`.

The reference moves the whole D x H x W DPV to the host and reduces it there
(/root/reference/code/test_utils/export_res.py:37-75, :77-100). Here one kernel reduces it on the
device to the expected-depth and confidence maps, already scaled and truncated to uint16, so the
only device->host traffic is the two maps (1.2 MB instead of 78.6 MB at 640x480x64); the .pgm files
are byte-identical to the reference's (PIL 'I' -> P5, 65535, big-endian).

Same names and argument order as the reference:
    depth_regression(Depth_Indx_vol, BV)
    export_res_img(ref_dat, BV_measure, d_candi, resfldr, batch_idx, depth_scale=1000, conf_scale=1000)
    export_res_refineNet(ref_dat, BV_measure, d_candi, res_fldr, batch_idx, ...)
The matplotlib previews of export_res_refineNet (input.png, conf.png, dmap_raw.png, dmaps_diff.png, dmap_ref.png and
their concatenation res_%05d.png, :104-141) are written by `_previews` when matplotlib is importable, exactly as the reference
needs it; without matplotlib (this image) that group is omitted - the numeric products (.mat, .pgm, 16-bit PNGs) always are.
rgb_%05d.png has the reference's cv2 (BGR) byte order.
"""
import ctypes
import os

import numpy as np
import torch

from .. import _lib
from .._devcache import planes_tensor
from .._lib import check, ptr

_IMAGENET = {'mean': [0.485, 0.456, 0.406], 'std': [0.229, 0.224, 0.225]}      # export_res.py:27-28


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def depth_conf_maps(BV_measure, d_candi, depth_scale=1000, conf_scale=1000, want_float=True, want_u16=True):
    """Fused reduction of a log-DPV `1 x D x H x W` (CUDA) -> dict with any of
    dmap, conf (float32 H x W, device) and dmap_u16, conf_u16 (uint16 as int16-storage tensors, device)."""
    if not BV_measure.is_cuda:
        raise _lib.NrgbdError('neuralrgbd_b200 has no CPU path: expected a CUDA tensor')
    assert BV_measure.dim() == 4 and BV_measure.shape[0] == 1, 'BV_measure should be 1 x D x H x W'
    assert len(d_candi) == BV_measure.shape[1], 'BV_measure should have the same # of slices as len(d_candi) !'
    D, H, W = BV_measure.shape[1:]
    bv = BV_measure.detach().contiguous().float()
    dev = bv.device
    dc = planes_tensor(d_candi, dev)
    out = {}
    if want_float:
        out['dmap'] = torch.empty((H, W), device=dev, dtype=torch.float32)
        out['conf'] = torch.empty((H, W), device=dev, dtype=torch.float32)
    if want_u16:
        out['dmap_u16'] = torch.empty((H, W), device=dev, dtype=torch.int16)     # raw uint16 bits
        out['conf_u16'] = torch.empty((H, W), device=dev, dtype=torch.int16)
    L = _lib.lib()
    check(L.nrgbd_export_depth_conf(ptr(bv), ptr(dc), D, H * W, ctypes.c_float(depth_scale), ctypes.c_float(conf_scale),
                                    ptr(out.get('dmap')), ptr(out.get('conf')), ptr(out.get('dmap_u16')), ptr(out.get('conf_u16')),
                                    _stream()))
    return out


def _u16_host(t):
    return t.cpu().numpy().view(np.uint16)


def write_pgm16(path, im_u16):
    """mio/imgIO.py:9-10 export2pgm for a uint16 H x W array (host)."""
    im = np.ascontiguousarray(im_u16, dtype=np.uint16)
    check(_lib.lib().nrgbd_write_pgm16(os.fsencode(path), ctypes.c_void_p(im.ctypes.data), im.shape[1], im.shape[0]))


def depth_regression(Depth_Indx_vol, BV):
    """export_res.py:37-41: sum_d exp(BV) * Depth_Indx_vol -> numpy H x W. Depth_Indx_vol is the constant-per-plane
    volume the reference builds (:49-52); only its per-plane values are used here."""
    d = Depth_Indx_vol.reshape(Depth_Indx_vol.shape[-3], -1)[:, 0].double().cpu().numpy()
    return depth_conf_maps(BV, d, want_u16=False)['dmap'].cpu().numpy()


def _un_normalize(img_in):
    img_out = np.zeros(img_in.shape)
    for ich in range(3):
        img_out[:, :, ich] = img_in[:, :, ich] * _IMAGENET['std'][ich] + _IMAGENET['mean'][ich]
    return img_out


def _save_rgb(path, arr_u8):
    try:
        import PIL.Image as image
        image.fromarray(arr_u8).save(path)
    except ImportError:          # no image library: the numeric outputs are still written
        pass


def export_res_img(ref_dat, BV_measure, d_candi, resfldr, batch_idx, depth_scale=1000, conf_scale=1000):
    """export_res.py:43-75: writes img_%05d.png, d_%05d.pgm (depth * depth_scale, uint16) and
    conf_%05d.pgm (max probability * conf_scale, uint16)."""
    maps = depth_conf_maps(BV_measure, d_candi, depth_scale, conf_scale, want_float=False)
    os.makedirs(resfldr, exist_ok=True)
    img = ref_dat['img'].squeeze().cpu().permute(1, 2, 0).numpy()
    img_in_png = (_un_normalize(img) * 255).astype(np.uint8)
    _save_rgb('%s/img_%05d.png' % (resfldr, batch_idx), img_in_png)
    write_pgm16('%s/d_%05d.pgm' % (resfldr, batch_idx), _u16_host(maps['dmap_u16']))
    write_pgm16('%s/conf_%05d.pgm' % (resfldr, batch_idx), _u16_host(maps['conf_u16']))


def _previews(resfldr, batch_idx, img_in_raw, conf, dmap, dmap_ref, d_max, diff_vrange_ratio):
    """The colour-mapped preview files of export_res.py:104-141 (input.png, conf.png, dmap_raw.png, [dmaps_diff.png,
    dmap_ref.png] and their horizontal concatenation res_%05d.png). They need matplotlib, as in the reference; when it is not
    importable NOTHING of this group is written (the numeric products - .mat, .pgm, 16-bit PNGs - do not depend on it)."""
    try:
        import matplotlib as mlt
        mlt.use('Agg')
        import matplotlib.pyplot as plt
        import PIL.Image as image
    except ImportError:
        return False
    img_in = (img_in_raw - img_in_raw.min()) / (img_in_raw.max() - img_in_raw.min()) * 255.
    names = ['%s/input.png' % resfldr, '%s/conf.png' % resfldr, '%s/dmap_raw.png' % resfldr]
    if dmap_ref is not None:
        mask = (dmap_ref > 0).astype(np.float64)
        plt.imsave('%s/dmaps_diff.png' % resfldr, np.abs(dmap_ref - dmap) * mask, vmin=0, vmax=d_max / diff_vrange_ratio)
        plt.imsave('%s/dmap_ref.png' % resfldr, dmap_ref, vmax=d_max, vmin=0, cmap='gray')
        names += ['%s/dmaps_diff.png' % resfldr, '%s/dmap_ref.png' % resfldr]
    plt.imsave(names[1], conf, vmin=0, vmax=1, cmap='jet')
    plt.imsave(names[2], dmap, vmin=0., vmax=d_max, cmap='gray')
    plt.imsave(names[0], img_in.astype(np.uint8))
    plt.imsave('%s/res_%05d.png' % (resfldr, batch_idx), np.hstack([np.array(image.open(n)) for n in names]))      # cat_imgs :30-33
    return True


def export_res_refineNet(ref_dat, BV_measure, d_candi, res_fldr, batch_idx, diff_vrange_ratio=4,
                         cam_pose=None, cam_intrinM=None, output_pngs=False, save_mat=True, output_dmap_ref=True):
    """export_res.py:77-160: depth / confidence maps of the refined DPV, .mat dump and optional 16-bit PNGs."""
    maps = depth_conf_maps(BV_measure, d_candi, 1000, 255)
    dmap = maps['dmap'].cpu().numpy()
    confMap_log = maps['conf'].cpu().numpy()
    img_in_raw = ref_dat['img'].squeeze().cpu().permute(1, 2, 0).numpy()
    os.makedirs(res_fldr, exist_ok=True)
    dmap_ref = None
    if output_dmap_ref:
        dmap_ref = ref_dat['dmap_imgsize'].squeeze().cpu().numpy()
    _previews(res_fldr, batch_idx, img_in_raw, confMap_log, dmap, dmap_ref, float(np.max(d_candi)), diff_vrange_ratio)
    if save_mat:
        import scipy.io as sio
        mdict = {'dmap': dmap, 'img': img_in_raw, 'confMap': confMap_log, 'img_path': ref_dat['img_path']}      # KeyError as in the reference (:119)
        if output_dmap_ref:
            mdict['dmap_ref'] = dmap_ref
            if cam_pose is not None:
                mdict['cam_pose'] = cam_pose
                mdict['cam_intrinM'] = cam_intrinM
        sio.savemat('%s/depth_%05d.mat' % (res_fldr, batch_idx), mdict)
    if output_pngs:
        png_fldr = '%s/output_pngs' % (res_fldr,)
        os.makedirs(png_fldr, exist_ok=True)
        try:
            import PIL.Image as image
            image.fromarray(_u16_host(maps['dmap_u16'])).save('%s/d_%05d.png' % (png_fldr, batch_idx))
            # the reference writes this RGB array with cv2.imwrite, which stores it as if it were BGR (:152): same file contents here
            image.fromarray(np.ascontiguousarray((_un_normalize(img_in_raw) * 255).astype(np.uint8)[:, :, ::-1])).save('%s/rgb_%05d.png' % (png_fldr, batch_idx))
            image.fromarray(_u16_host(maps['conf_u16']).astype(np.uint8)).save('%s/conf_%05d.png' % (png_fldr, batch_idx))
            if output_dmap_ref:
                image.fromarray((dmap_ref * 1000).astype(np.uint16)).save('%s/dref_%05d.png' % (png_fldr, batch_idx))
        except ImportError:
            pass
    return dmap, confMap_log
