// The zoom-video side of a headless viewer (reference viewer.cpp:271-285 drives video.cpp:8-34): key frames coloured at
// 1.5x the video size go into VideoZoom::nextFrame, `rate` in-between canvases per pair come out of the writer.
// Compiled against include/newman_b200/video.h, linked with libnewman_b200.so only (tests/test_dropin_cpp.py).
//   headless_video OUT NR NC RATE KEYS     key frame k: pixel (r, c, ch) = (7 r + 13 c + 29 ch + 101 k + r c k) mod 256
#include "video.h"

#include <cstdio>
#include <cstdlib>
#include <stdexcept>

int main(int argc, char** argv) {
  if (argc < 6) { fprintf(stderr, "usage: headless_video OUT NR NC RATE KEYS\n"); return 1; }
  const int nr = atoi(argv[2]), nc = atoi(argv[3]), rate = atoi(argv[4]), keys = atoi(argv[5]);
  try {
    VideoZoom zoom;
    zoom.start(argv[1], nr, nc, rate);
    for (int k = 0; k < keys; k++) {
      ByteImage img(nr * 3 / 2, nc * 3 / 2, 3);   // viewer.cpp:272
      for (int r = 0; r < img.nr; r++)
        for (int c = 0; c < img.nc; c++)
          for (int ch = 0; ch < 3; ch++) img.at(r, c, ch) = (unsigned char)((7 * r + 13 * c + 29 * ch + 101 * k + r * c * k) & 255);
      zoom.nextFrame(img);
      VideoZoom copy = zoom;   // a value type like the reference's
      (void)copy;
    }
  } catch (const std::runtime_error& e) {
    fprintf(stderr, "runtime_error: %s\n", e.what());
    return 2;
  }
  printf("ok\n");
  return 0;
}
