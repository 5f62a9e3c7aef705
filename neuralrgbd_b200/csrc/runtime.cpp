// Error string + launch counter shared by all nrgbd translation units.
#include <atomic>
#include <cstdarg>
#include <cstdio>

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void nrgbd_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" {
const char* nrgbd_last_error(void) { return g_err; }
void nrgbd_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long nrgbd_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
void nrgbd_reset_launch_count(void) { g_launches.store(0, std::memory_order_relaxed); }
int nrgbd_abi_version(void) { return 1; }

// 16-bit binary PGM exactly as PIL writes a mode-'I' image to a .pgm path (mio/imgIO.py:9-10 via
// test_utils/export_res.py:74-75): "P5\n<W> <H>\n65535\n" then big-endian samples. Host-only.
int nrgbd_write_pgm16(const char* path, const unsigned short* pixels, int width, int height) {
  if (!path || !pixels || width < 1 || height < 1) { nrgbd_set_error("nrgbd_write_pgm16: bad arguments"); return -1; }
  FILE* f = fopen(path, "wb");
  if (!f) { nrgbd_set_error("nrgbd_write_pgm16: cannot open %s", path); return -5; }
  fprintf(f, "P5\n%d %d\n65535\n", width, height);
  const size_t n = (size_t)width * (size_t)height;
  unsigned char buf[8192];
  size_t done = 0;
  bool ok = true;
  while (done < n && ok) {
    size_t m = n - done; if (m > sizeof(buf) / 2) m = sizeof(buf) / 2;
    for (size_t i = 0; i < m; ++i) { buf[2 * i] = (unsigned char)(pixels[done + i] >> 8); buf[2 * i + 1] = (unsigned char)(pixels[done + i] & 0xff); }
    ok = fwrite(buf, 2, m, f) == m;
    done += m;
  }
  ok = (fclose(f) == 0) && ok;
  if (!ok) { nrgbd_set_error("nrgbd_write_pgm16: short write to %s", path); return -5; }
  return 0;
}
}
