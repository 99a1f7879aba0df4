// WAV in / out: the steps either side of the render path, host-side and renderer-free.
//   in : what AudioFileBuffer::from_file yields for a RIFF/WAVE file (src/source/file/buffer.rs:64-119,
//        src/source/file/decoder.rs:170-330): interleaved f32 samples, channel count, rate and the first `smpl` loop.
//        The reference decodes through symphonia ^0.5 (third-party, not vendored): integer PCM is scaled by
//        1 / 2^(bits-1) (u8: (x - 128) / 128), float32 passes through -- the assumption SURVEY.md §8c flags.
//   out: the 32-bit float WAV WavOutput writes through hound ^3.5 (src/output/wav.rs:61-70): bits_per_sample 32,
//        SampleFormat::Float -> a WAVE_FORMAT_EXTENSIBLE `fmt ` chunk with the IEEE-float sub-format, samples as raw
//        little-endian f32 (the data bytes are what the reference writes; the header layout is hound's as documented).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/phonic_b200.h"

namespace pbh {

struct WavData {
  std::vector<float> samples;  // interleaved, without the pad frame
  pb200_wav_info info;
};

inline uint32_t rd_u32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline uint16_t rd_u16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }

inline int decode_wav_file(const char* path, WavData& out) {
  FILE* f = std::fopen(path, "rb");
  if (!f) return PB200_ERR_MEDIA_FILE_NOT_FOUND;
  std::vector<uint8_t> d;
  uint8_t tmp[1 << 16];
  size_t n;
  while ((n = std::fread(tmp, 1, sizeof(tmp), f)) > 0) d.insert(d.end(), tmp, tmp + n);
  std::fclose(f);
  if (d.size() < 12 || std::memcmp(d.data(), "RIFF", 4) != 0 || std::memcmp(d.data() + 8, "WAVE", 4) != 0) return PB200_ERR_MEDIA_FILE_PROBE;
  uint32_t tag = 0, channels = 0, rate = 0, bits = 0, block_align = 0;
  const uint8_t* data = nullptr;
  size_t data_len = 0;
  int64_t loop_start = PB200_NO_LOOP, loop_end = PB200_NO_LOOP;
  bool have_loop = false;
  for (size_t p = 12; p + 8 <= d.size();) {
    const uint8_t* ck = d.data() + p;
    const size_t sz = rd_u32(ck + 4);
    const size_t body = p + 8, avail = d.size() - body;
    const size_t len = sz < avail ? sz : avail;  // a truncated last chunk is read as far as it goes
    if (std::memcmp(ck, "fmt ", 4) == 0 && len >= 16) {
      tag = rd_u16(ck + 8); channels = rd_u16(ck + 10); rate = rd_u32(ck + 12); block_align = rd_u16(ck + 20); bits = rd_u16(ck + 22);
      if (tag == 0xFFFE && len >= 26) tag = rd_u16(ck + 8 + 24);  // WAVE_FORMAT_EXTENSIBLE: first word of the SubFormat GUID
    } else if (std::memcmp(ck, "data", 4) == 0 && !data) {
      data = ck + 8; data_len = len;
    } else if (std::memcmp(ck, "smpl", 4) == 0 && !have_loop && len >= 36) {
      // parse_smpl_body (decoder.rs:294-330): loop count at 28, 24-byte entries from 36, start / end at +8 / +12
      const uint32_t n_loops = rd_u32(ck + 8 + 28);
      if (n_loops > 0 && len >= 36 + 24) {
        loop_start = rd_u32(ck + 8 + 36 + 8);
        loop_end = rd_u32(ck + 8 + 36 + 12);
        have_loop = true;
      }
    }
    p = body + sz + (sz & 1);
  }
  if (!data || channels == 0 || rate == 0) return PB200_ERR_MEDIA_FILE_PROBE;
  const uint32_t bytes = bits / 8;
  const bool is_float = tag == 3;
  if (!((tag == 1 && (bits == 8 || bits == 16 || bits == 24 || bits == 32)) || (is_float && (bits == 32 || bits == 64)))) return PB200_ERR_MEDIA_FILE_PROBE;
  if (block_align != bytes * channels) return PB200_ERR_MEDIA_FILE_PROBE;
  const size_t frames = data_len / block_align;
  if (frames == 0) return PB200_ERR_AUDIO_DECODING;  // "failed to decode file" (buffer.rs:97-101)
  out.samples.resize(frames * channels);
  for (size_t i = 0; i < frames * channels; ++i) {
    const uint8_t* s = data + i * bytes;
    float v;
    if (is_float && bits == 32) { uint32_t u = rd_u32(s); std::memcpy(&v, &u, 4); }
    else if (is_float) { uint64_t u = (uint64_t)rd_u32(s) | ((uint64_t)rd_u32(s + 4) << 32); double dd; std::memcpy(&dd, &u, 8); v = (float)dd; }
    else if (bits == 8) v = (float)((int)s[0] - 128) / 128.0f;
    else if (bits == 16) v = (float)(int16_t)rd_u16(s) / 32768.0f;
    else if (bits == 24) { int32_t x = (int32_t)(s[0] | (s[1] << 8) | (s[2] << 16)); if (x & 0x800000) x |= ~0xFFFFFF; v = (float)x / 8388608.0f; }
    else v = (float)((double)(int32_t)rd_u32(s) / 2147483648.0);
    out.samples[i] = v;
  }
  std::memset(&out.info, 0, sizeof(out.info));
  out.info.frames = frames; out.info.channels = channels; out.info.sample_rate = rate;
  out.info.bits_per_sample = bits; out.info.is_float = is_float ? 1u : 0u;
  out.info.loop_start = PB200_NO_LOOP; out.info.loop_end = PB200_NO_LOOP;
  if (have_loop) {  // buffer.rs:105-115: clamped to the frame count INCLUDING the pad frame; kept only if end > start
    const int64_t fc = (int64_t)frames + 1;
    const int64_t ls = loop_start < fc ? loop_start : fc, le = loop_end < fc ? loop_end : fc;
    if (le > ls) { out.info.loop_start = ls; out.info.loop_end = le; }
  }
  return PB200_OK;
}

inline void put_u32(std::vector<uint8_t>& b, uint32_t v) { for (int i = 0; i < 4; ++i) b.push_back((uint8_t)(v >> (8 * i))); }
inline void put_u16(std::vector<uint8_t>& b, uint16_t v) { b.push_back((uint8_t)v); b.push_back((uint8_t)(v >> 8)); }

inline int write_wav_f32(const char* path, const float* interleaved, uint64_t frames, uint32_t channels, uint32_t rate) {
  const uint64_t data_bytes = frames * channels * 4ull;
  if (data_bytes > 0xFFFFFFFFull - 68) return PB200_ERR_IO;  // RIFF sizes are 32 bit
  std::vector<uint8_t> h;
  h.insert(h.end(), {'R', 'I', 'F', 'F'}); put_u32(h, (uint32_t)(4 + 8 + 40 + 8 + data_bytes));
  h.insert(h.end(), {'W', 'A', 'V', 'E', 'f', 'm', 't', ' '}); put_u32(h, 40);
  put_u16(h, 0xFFFE); put_u16(h, (uint16_t)channels); put_u32(h, rate); put_u32(h, rate * channels * 4); put_u16(h, (uint16_t)(channels * 4));
  put_u16(h, 32); put_u16(h, 22); put_u16(h, 32);
  put_u32(h, channels == 1 ? 0x4u : (channels == 2 ? 0x3u : 0u));  // speaker mask: mono = front centre, stereo = front left | right
  static const uint8_t ieee_float_guid[16] = {0x03, 0x00, 0x00, 0x00, 0x00, 0x00, 0x10, 0x00, 0x80, 0x00, 0x00, 0xAA, 0x00, 0x38, 0x9B, 0x71};
  h.insert(h.end(), ieee_float_guid, ieee_float_guid + 16);
  h.insert(h.end(), {'d', 'a', 't', 'a'}); put_u32(h, (uint32_t)data_bytes);
  FILE* f = std::fopen(path, "wb");
  if (!f) return PB200_ERR_IO;
  bool ok = std::fwrite(h.data(), 1, h.size(), f) == h.size();
  ok = ok && std::fwrite(interleaved, 4, frames * channels, f) == frames * channels;  // little-endian host
  ok = (std::fclose(f) == 0) && ok;
  return ok ? PB200_OK : PB200_ERR_IO;
}

// WavStream::process (wav.rs:210-250): whole blocks while whole-seconds(pos / rate) < duration
inline uint64_t wav_stream_frames(uint64_t duration_nanos, uint32_t rate, uint32_t block_frames) {
  uint64_t frames = 0;
  while ((frames / rate) * 1000000000ull < duration_nanos) frames += block_frames;
  return frames;
}

}  // namespace pbh
