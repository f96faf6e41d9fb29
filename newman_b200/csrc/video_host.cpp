// Host side of the zoom-video drop-in (include/newman_b200/video.h == reference video.h:13-26): VideoZoom::start /
// nextFrame (video.cpp:8-34) over K5. The frame loop of the reference becomes one nm_video_inbetween call per key frame.
#include "../../include/newman_b200/video.h"

#include <stdexcept>

#include "../../include/newman_b200.h"

#ifndef NEWMAN_B200_HAVE_BYTEIMAGE
namespace byteimage {

void VideoWriter::open(const std::string& name, int nr, int nc, int /*fps*/) {
  nr_ = nr; nc_ = nc;
  FILE* f = fopen(name.c_str(), "wb");
  fp_ = f ? std::shared_ptr<FILE>(f, fclose) : std::shared_ptr<FILE>();   // like the reference: silent on I/O failure
}

void VideoWriter::write(const ByteImage& frame) {
  if (!fp_ || frame.nr != nr_ || frame.nc != nc_ || frame.nchannels != 3) return;
  std::vector<unsigned char> row((size_t)nc_ * 3);
  for (int r = 0; r < nr_; r++) {
    for (int c = 0; c < nc_; c++)
      for (int ch = 0; ch < 3; ch++) row[(size_t)c * 3 + ch] = frame.at(r, c, ch);
    fwrite(row.data(), 1, row.size(), fp_.get());
  }
}

}  // namespace byteimage
#endif

namespace {

// (r, c, ch) bytes, interleaved, of a 3-channel image (a grey one is replicated)
std::vector<unsigned char> interleaved(const ByteImage& im) {
  std::vector<unsigned char> out((size_t)im.nr * im.nc * 3);
  for (int r = 0; r < im.nr; r++)
    for (int c = 0; c < im.nc; c++)
      for (int ch = 0; ch < 3; ch++) out[((size_t)r * im.nc + c) * 3 + ch] = im.at(r, c, im.nchannels >= 3 ? ch : 0);
  return out;
}

}  // namespace

VideoZoom::VideoZoom() : nr(0), nc(0), rate(30) {}

void VideoZoom::start(const std::string& name, int nr_, int nc_, int rate_) {
  writer.open(name, nr_, nc_, 30);
  img = ByteImage();
  nr = nr_;
  nc = nc_;
  rate = rate_;
}

void VideoZoom::nextFrame(const ByteImage& next) {
  if (img.size()) {
    if (next.nr != img.nr || next.nc != img.nc) throw std::runtime_error("VideoZoom::nextFrame: key frames differ in size");
    if (nr < 1 || nc < 1 || rate < 1) throw std::runtime_error("VideoZoom::nextFrame: start() was not called");
    if (!gpu_) {
      nm_ctx* ctx = nullptr;
      const int rc = nm_create(0, &ctx);
      if (rc != NM_OK) throw std::runtime_error(std::string("VideoZoom: ") + nm_last_error(nullptr));
      gpu_ = std::shared_ptr<void>(ctx, [](void* p) { nm_destroy((nm_ctx*)p); });
    }
    nm_ctx* ctx = (nm_ctx*)gpu_.get();
    const std::vector<unsigned char> prev_rgb = interleaved(img), next_rgb = interleaved(next);
    std::vector<unsigned char> frames((size_t)rate * nr * nc * 3);
    const int rc = nm_video_inbetween(ctx, prev_rgb.data(), next_rgb.data(), img.nr, img.nc, nr, nc, rate, frames.data());
    if (rc != NM_OK) throw std::runtime_error(std::string("VideoZoom::nextFrame: ") + nm_last_error(ctx));
    ByteImage canvas(nr, nc, 3);
    for (int i = 0; i < rate; i++) {
      const unsigned char* f = frames.data() + (size_t)i * nr * nc * 3;
      for (int r = 0; r < nr; r++)
        for (int c = 0; c < nc; c++)
          for (int ch = 0; ch < 3; ch++) canvas.at(r, c, ch) = f[((size_t)r * nc + c) * 3 + ch];
      writer.write(canvas);
    }
  }
  img = next;
}
