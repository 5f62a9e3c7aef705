/* Minimal C consumer of the libnrgbd C ABI (no Python, no torch, no CUDA headers): writes a 16-bit depth map the
 * way the reference's export does (mio/imgIO.py:9-10, test_utils/export_res.py:74) through nrgbd_write_pgm16.
 *
 *   gcc -Iinclude examples/write_depth_pgm.c -Lneuralrgbd_b200 -lnrgbd -Wl,-rpath,$PWD/neuralrgbd_b200 -o write_depth_pgm
 *   ./write_depth_pgm out.pgm 64 48
 *
 * A device-side caller would first fill the uint16 maps with nrgbd_export_depth_conf. */
#include <stdio.h>
#include <stdlib.h>

#include "nrgbd.h"

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: %s out.pgm width height\n", argv[0]); return 2; }
  const int w = atoi(argv[2]), h = atoi(argv[3]);
  if (w < 1 || h < 1) { fprintf(stderr, "bad size\n"); return 2; }
  unsigned short* depth_mm = (unsigned short*)malloc(sizeof(unsigned short) * (size_t)w * (size_t)h);
  if (!depth_mm) return 1;
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) depth_mm[(size_t)y * w + x] = (unsigned short)(500 + 7 * x + 3 * y);   /* a ramp, in mm */
  const int rc = nrgbd_write_pgm16(argv[1], depth_mm, w, h);
  if (rc != NRGBD_OK) fprintf(stderr, "nrgbd_write_pgm16 failed (%d): %s\n", rc, nrgbd_last_error());
  else printf("libnrgbd ABI %d: wrote %dx%d 16-bit PGM to %s\n", nrgbd_abi_version(), w, h, argv[1]);
  free(depth_mm);
  return rc == NRGBD_OK ? 0 : 1;
}
