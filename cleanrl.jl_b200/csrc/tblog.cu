// tblog.cu — native TensorBoard event-file writer behind the logger hook (logger.jl:7-29 builds a TBLogger; the PPO loop
// emits one "Episode Statistics" and num_minibatches x update_epochs "Training Statistics" records per update,
// ppo.jl:157,247). Host-only code. At the reference's default shape (4 envs x 32 steps) an update takes a few hundred
// microseconds on the GPU, and building one protobuf per scalar in Python (~100 us each) made the LOGGER the bottleneck
// of ppo(PPOConfig()); this writer serialises a record in well under a microsecond.
//
// File format (TensorBoard / TFRecord): every record is  u64 length | u32 masked_crc32c(length) | data | u32
// masked_crc32c(data)  with data = a serialised tensorflow.Event: wall_time (field 1, double), step (field 2, int64),
// file_version (field 3, string, first record only) or summary (field 5) holding one Summary.Value {tag (1, string),
// simple_value (2, float)} per scalar. All scalars of one record share one Event, as TensorBoardLogger.jl writes them.
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include <string>
#include <vector>

#include "../../include/cleanrl_cuda.h"

int crl_internal_fail(int code, const char* msg);

struct crl_tb {
  FILE* f;
  std::string buf;
};

namespace {

uint32_t g_crc_table[256];
bool g_crc_ready = false;
void crc_init() {
  for (uint32_t i = 0; i < 256; i++) {
    uint32_t c = i;
    for (int k = 0; k < 8; k++) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;  // CRC-32C (Castagnoli), reflected
    g_crc_table[i] = c;
  }
  g_crc_ready = true;
}
uint32_t crc32c(const unsigned char* p, size_t n) {
  uint32_t c = 0xFFFFFFFFu;
  for (size_t i = 0; i < n; i++) c = g_crc_table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
  return c ^ 0xFFFFFFFFu;
}
uint32_t masked_crc(const unsigned char* p, size_t n) {
  const uint32_t c = crc32c(p, n);
  return ((c >> 15) | (c << 17)) + 0xA282EAD8u;
}
void put_varint(std::string& s, uint64_t v) {
  while (v >= 0x80) { s.push_back((char)((v & 0x7F) | 0x80)); v >>= 7; }
  s.push_back((char)v);
}
void put_bytes(std::string& s, const void* p, size_t n) { s.append(reinterpret_cast<const char*>(p), n); }

void write_record(crl_tb* t, const std::string& data) {
  unsigned char head[12];
  const uint64_t len = data.size();
  memcpy(head, &len, 8);
  const uint32_t c1 = masked_crc(head, 8);
  memcpy(head + 8, &c1, 4);
  const uint32_t c2 = masked_crc(reinterpret_cast<const unsigned char*>(data.data()), data.size());
  fwrite(head, 1, 12, t->f);
  fwrite(data.data(), 1, data.size(), t->f);
  fwrite(&c2, 1, 4, t->f);
}

}  // namespace

extern "C" CRL_API int crl_tb_open(const char* logdir, crl_tb** out) {
  if (!logdir || !out) return crl_internal_fail(CRL_ERR_INVALID, "crl_tb_open: NULL argument");
  if (!g_crc_ready) crc_init();
  char host[256] = "localhost";
  gethostname(host, sizeof(host) - 1);
  char path[4096];
  const double now = (double)time(nullptr);
  snprintf(path, sizeof(path), "%s/events.out.tfevents.%010.0f.%s.%d.crl", logdir, now, host, (int)getpid());
  FILE* f = fopen(path, "wb");
  if (!f) return crl_internal_fail(CRL_ERR_INVALID, "crl_tb_open: cannot create the event file (does the directory exist?)");
  setvbuf(f, nullptr, _IOFBF, 1 << 20);
  crl_tb* t = new crl_tb();
  t->f = f;
  std::string& d = t->buf;
  d.clear();
  d.push_back(0x09); put_bytes(d, &now, 8);                        // wall_time
  const char* ver = "brain.Event:2";
  d.push_back(0x1a); put_varint(d, strlen(ver)); put_bytes(d, ver, strlen(ver));  // file_version
  write_record(t, d);
  *out = t;
  return CRL_OK;
}

// tags: n NUL-terminated strings back to back ("Training Statistics/loss\0Training Statistics/pg_loss\0...")
extern "C" CRL_API int crl_tb_scalars(crl_tb* t, double wall_time, int64_t step, int32_t n, const char* tags, const double* values) {
  if (!t || (n > 0 && (!tags || !values))) return crl_internal_fail(CRL_ERR_INVALID, "crl_tb_scalars: NULL argument");
  std::string summary;
  const char* tag = tags;
  for (int i = 0; i < n; i++) {
    const size_t tl = strlen(tag);
    std::string val;
    val.push_back(0x0a); put_varint(val, tl); put_bytes(val, tag, tl);     // Value.tag
    const float fv = (float)values[i];
    val.push_back(0x15); put_bytes(val, &fv, 4);                           // Value.simple_value
    summary.push_back(0x0a); put_varint(summary, val.size()); summary += val;  // Summary.value
    tag += tl + 1;
  }
  std::string& d = t->buf;
  d.clear();
  d.push_back(0x09); put_bytes(d, &wall_time, 8);
  d.push_back(0x10); put_varint(d, (uint64_t)step);
  d.push_back(0x2a); put_varint(d, summary.size()); d += summary;
  write_record(t, d);
  return CRL_OK;
}

extern "C" CRL_API int crl_tb_flush(crl_tb* t) {
  if (!t) return crl_internal_fail(CRL_ERR_INVALID, "crl_tb_flush: NULL handle");
  fflush(t->f);
  return CRL_OK;
}

extern "C" CRL_API int crl_tb_close(crl_tb* t) {
  if (!t) return CRL_OK;
  fclose(t->f);
  delete t;
  return CRL_OK;
}
