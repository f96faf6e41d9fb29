// newman_b200/video.h — drop-in for newman's zoom-video writer (reference video.h:13-26, video.cpp:3-34).
//
// VideoZoom::nextFrame receives the key frames of a zoom (each rendered 1.5x deeper than the one before and coloured
// at 1.5x the video size, viewer.cpp:271-285) and writes `rate` in-between canvases per key frame. The reference does
// that with libbyteimage's scaled / blit / blend on the host, one canvas at a time (video.cpp:17-31); here all `rate`
// canvases of a key-frame pair come from ONE launch of K5 (nm_video_inbetween in <newman_b200.h>, csrc/k5_video.cuh)
// and are handed to the writer in order. Same class, same members, same calls.
//
// libbyteimage is not vendored by the reference (README.md:24). With it installed define NEWMAN_B200_HAVE_BYTEIMAGE and
// its ByteImage / VideoWriter are used. Otherwise the minimal stand-ins below apply: a planar byte image with the
// members video.cpp and viewer.cpp touch, and a writer that appends raw RGB24 frames to the file (readable with
// `ffmpeg -f rawvideo -pix_fmt rgb24 -s WxH -r 30 -i file`): encoding is libbyteimage's business, not this path's.
// Without a CUDA device nextFrame throws std::runtime_error (no CPU fallback).
#ifndef NEWMAN_B200_VIDEO_H
#define NEWMAN_B200_VIDEO_H

#include <cstdio>
#include <memory>
#include <string>
#include <vector>

#ifndef NEWMAN_B200_CXX_API
#define NEWMAN_B200_CXX_API __attribute__((visibility("default")))   // exported from libnewman_b200.so (see mandelbrot.h)
#endif

#ifdef NEWMAN_B200_HAVE_BYTEIMAGE
#include <byteimage/video.h>
#else
namespace byteimage {
class ByteImage {   // planar: channel ch of pixel (r, c) at pixels[(ch * nr + r) * nc + c]
public:
  int nr, nc, nchannels;
  std::vector<unsigned char> pixels;
  ByteImage() : nr(0), nc(0), nchannels(0) {}
  ByteImage(int rows, int cols, int channels = 1)
      : nr(rows), nc(cols), nchannels(channels), pixels((size_t)rows * cols * channels) {}
  size_t size() const { return pixels.size(); }
  unsigned char& at(int r, int c, int ch = 0) { return pixels[((size_t)ch * nr + r) * nc + c]; }
  const unsigned char& at(int r, int c, int ch = 0) const { return pixels[((size_t)ch * nr + r) * nc + c]; }
};
class NEWMAN_B200_CXX_API VideoWriter {   // raw RGB24 frames, interleaved, appended to the file
  std::shared_ptr<FILE> fp_;
  int nr_, nc_;
public:
  VideoWriter() : nr_(0), nc_(0) {}
  void open(const std::string& name, int nr, int nc, int fps);
  void write(const ByteImage& frame);
  void close() { fp_.reset(); }
};
}  // namespace byteimage
#endif

using byteimage::ByteImage;
using byteimage::VideoWriter;

/*
 * Constraint inherited from the reference (video.h:9-11): consecutive key frames are 1.5x apart.
 */
class NEWMAN_B200_CXX_API VideoZoom {
protected:
  VideoWriter writer;
  ByteImage img;   // the previous key frame
  int nr, nc;
  int rate;        // frames to interpolate per zoom
  std::shared_ptr<void> gpu_;   // nm_ctx, created on first use, shared between copies

public:
  VideoZoom();

  void start(const std::string& name, int nr, int nc, int rate);
  void nextFrame(const ByteImage& img);
};

#endif
