// rmgpu_server -- persistent per-box worker behind the same C ABI (SURVEY 7 "hard part 5", 8f.4).
//
//   rmgpu_server --socket /tmp/rmgpu.sock [--gpus all|0,1,2,...] [--max-batch N]
//
// RepeatMasker builds its engine object in the parent and forks one short-lived child per batch
// (RepeatMasker:602, :1040-1180); a CUDA context cannot cross fork and creating one per child (plus re-reading and
// re-packing the library on every search(), as the reference does with cross_match) would dwarf the search itself.
// With RMGPU_SERVER=<socket> in the environment, rmgpu_search() in the children becomes a thin client: it sends the
// three paths and the parameters; this process runs the search on one of its GPU workers and returns the flat hit
// records.
//
//   * One worker thread per GPU with a queue.  A connection goes to the GPU with the shortest queue (per-GPU batch
//     queues; workers never talk to each other: no collective).
//   * COALESCING: a worker takes the oldest pending call and every other pending call with the same library, matrix and
//     parameters (-pa children of one stage differ only in their batch file) and searches all their batch files as ONE
//     batch -- one score pass per GPU instead of one per child -- then hands every caller the hits of its own records.
//     Records are independent in every step of the engine ('*' and masklevel are decided per query record), so the
//     results are those of the separate calls.
//   * Libraries stay resident per (path, size, mtime): a small LRU of sessions per GPU, so a run that alternates
//     libraries between stages does not re-pack them; residue codes are matrix-independent, so a matrix change (GC bins
//     per batch, RepeatMasker:4784-4821) only rebuilds the 8-byte row profiles on the device.
#include <signal.h>
#include <sys/prctl.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/un.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rmgpu.h"
#include "remote.hpp"

using namespace rmgpu;

namespace {

struct Request {
  rmgpu_params p;
  std::string query, subject, matrix, libKey, matKey;
  // filled by the worker
  std::mutex mu;
  std::condition_variable cv;
  bool done = false;
  int rc = 0;
  std::string err;
  std::vector<uint8_t> body;
};

struct Resident {
  std::string libKey, matKey;
  rmgpu_session* session = nullptr;
  long lastUse = 0;
};

struct Worker {
  int device = 0;
  std::mutex mu;
  std::condition_variable cv;
  std::deque<std::shared_ptr<Request>> queue;
  std::vector<Resident> resident;  // LRU of libraries on this GPU
  long served = 0, passes = 0, clock = 0;
  int inflight = 0;  // calls of the pass that is running (guarded by mu)
  std::thread thread;
};

std::vector<Worker*> g_workers;
std::atomic<bool> g_quit{false};
std::atomic<int> g_connections{0};
std::string g_socket;
int g_listen = -1;
int g_maxBatch = 32;
int g_maxResident = 4;
double g_lingerMs = 0.0;
FILE* g_trace = nullptr;  // --trace FILE: one line per pass (unix time of start and end, GPU, calls, records, hits)
std::mutex g_traceMu;

double unix_now() {
  return std::chrono::duration<double>(std::chrono::system_clock::now().time_since_epoch()).count();
}

std::string file_key(const std::string& path) {
  struct stat st;
  if (stat(path.c_str(), &st) != 0) return path + "|missing";
  return path + "|" + std::to_string((long long)st.st_size) + "|" + std::to_string((long long)st.st_mtim.tv_sec) + "." +
         std::to_string((long long)st.st_mtim.tv_nsec);
}

bool read_str(int fd, std::string& s) {
  uint32_t n = 0;
  if (!read_full(fd, &n, 4)) return false;
  if (n == 0xFFFFFFFFu) {
    s.clear();
    return true;
  }
  if (n > (1u << 20)) return false;
  s.resize(n);
  return n == 0 || read_full(fd, &s[0], n);
}

void reply_error(int fd, int status, const std::string& msg) {
  int32_t st = status ? status : 1;
  uint32_t n = (uint32_t)msg.size();
  write_full(fd, &st, 4);
  write_full(fd, &n, 4);
  write_full(fd, msg.data(), n);
}

bool same_search(const Request& a, const Request& b) {
  rmgpu_params x = a.p, y = b.p;
  x.device = y.device = 0;
  return a.libKey == b.libKey && a.matKey == b.matKey && memcmp(&x, &y, sizeof(x)) == 0;
}

// One pass: the batch files of `group` searched as one batch on this worker's GPU.
void run_group(Worker* w, std::vector<std::shared_ptr<Request>>& group) {
  Request& first = *group[0];
  int rc = 0;
  std::string err;
  const double t0 = unix_now();
  Resident* res = nullptr;
  for (auto& r : w->resident)
    if (r.libKey == first.libKey) res = &r;
  if (!res) {
    if ((int)w->resident.size() >= g_maxResident) {  // evict the least recently used library
      size_t lru = 0;
      for (size_t i = 1; i < w->resident.size(); i++)
        if (w->resident[i].lastUse < w->resident[lru].lastUse) lru = i;
      rmgpu_session_close(w->resident[lru].session);
      w->resident.erase(w->resident.begin() + (long)lru);
    }
    Resident r;
    r.session = rmgpu_session_open(w->device);
    if (!r.session) {
      rc = 1;
      err = rmgpu_last_error();
    } else {
      w->resident.push_back(r);
      res = &w->resident.back();
    }
  }
  rmgpu_result* out = nullptr;
  std::vector<int32_t> firstRecord(group.size() + 1, 0);
  if (!rc) {
    res->lastUse = ++w->clock;
    if (res->matKey != first.matKey) {
      rc = rmgpu_session_set_matrix_file(res->session, first.matrix.c_str());
      res->matKey = rc ? "" : first.matKey;
    }
    if (!rc && res->libKey != first.libKey) {
      rc = rmgpu_session_set_library_fasta(res->session, first.subject.c_str());
      res->libKey = rc ? "" : first.libKey;
    }
    std::vector<const char*> files;
    for (auto& r : group) files.push_back(r->query.c_str());
    if (!rc) rc = rmgpu_session_set_queries_fasta_multi(res->session, (int32_t)files.size(), files.data(), firstRecord.data());
    rmgpu_params p = first.p;
    p.device = w->device;
    if (!rc) rc = rmgpu_session_search(res->session, &p, &out);
    if (rc && err.empty()) err = rmgpu_last_error();
  }
  // every caller gets the hits of its own records (hits come ordered by query record)
  int at = 0;
  for (size_t g = 0; g < group.size(); g++) {
    Request& r = *group[g];
    if (!rc) {
      rmgpu_result part = *out;
      int b = at;
      while (b < out->n_hits && out->hits[b].query_index < firstRecord[g + 1]) b++;
      std::vector<rmgpu_hit> hits(out->hits + at, out->hits + b);
      for (auto& h : hits) h.query_index -= firstRecord[g];
      part.n_hits = (int32_t)hits.size();
      part.hits = hits.data();
      part.opaque = nullptr;
      serialize_result(&part, r.body);
      at = b;
    }
    std::lock_guard<std::mutex> lk(r.mu);
    r.rc = rc;
    r.err = err;
    r.done = true;
    r.cv.notify_all();
  }
  if (g_trace) {
    std::lock_guard<std::mutex> lk(g_traceMu);
    fprintf(g_trace, "%.3f %.3f gpu %d calls %zu records %d hits %d rc %d\n", t0, unix_now(), w->device, group.size(),
            (int)firstRecord[group.size()], out ? out->n_hits : 0, rc);
    fflush(g_trace);
  }
  if (out) rmgpu_free(out);
  w->served += (long)group.size();
  w->passes++;
}

void worker_loop(Worker* w) {
  for (;;) {
    std::vector<std::shared_ptr<Request>> group;
    {
      std::unique_lock<std::mutex> lk(w->mu);
      w->cv.wait(lk, [&] { return !w->queue.empty() || g_quit.load(); });
      if (w->queue.empty()) break;
      if (g_lingerMs > 0) {  // optional: let the other children of the same stage arrive
        lk.unlock();
        std::this_thread::sleep_for(std::chrono::duration<double, std::milli>(g_lingerMs));
        lk.lock();
      }
      group.push_back(w->queue.front());
      w->queue.pop_front();
      for (auto it = w->queue.begin(); it != w->queue.end() && (int)group.size() < g_maxBatch;) {
        if (same_search(*group[0], **it)) {
          group.push_back(*it);
          it = w->queue.erase(it);
        } else
          ++it;
      }
      w->inflight = (int)group.size();
    }
    run_group(w, group);
    {
      std::lock_guard<std::mutex> lk(w->mu);
      w->inflight = 0;
    }
  }
}

void handle(int fd) {
  struct Guard {
    ~Guard() { g_connections--; }
  } guard;
  uint32_t magic = 0;
  if (!read_full(fd, &magic, 4)) {
    ::close(fd);
    return;
  }
  if (magic == kWirePing || magic == kWireQuit) {
    int32_t st = 0;
    write_full(fd, &st, 4);
    ::close(fd);
    if (magic == kWireQuit) {
      g_quit = true;
      ::shutdown(g_listen, SHUT_RDWR);
    }
    return;
  }
  auto req = std::make_shared<Request>();
  if (magic != kWireSearch || !read_full(fd, &req->p, sizeof(req->p)) || !read_str(fd, req->query) || !read_str(fd, req->subject) ||
      !read_str(fd, req->matrix)) {
    reply_error(fd, 1, "rmgpu_server: malformed request");
    ::close(fd);
    return;
  }
  req->libKey = file_key(req->subject);
  req->matKey = file_key(req->matrix);
  // the GPU with the least pending work (queued + in flight); the scan starts at a rotating index so that equally idle
  // GPUs take turns
  static std::atomic<unsigned> next{0};
  const unsigned start = next++ % (unsigned)g_workers.size();
  Worker* best = nullptr;
  size_t bestLen = 0;
  for (size_t i = 0; i < g_workers.size(); i++) {
    Worker* w = g_workers[(start + i) % g_workers.size()];
    std::lock_guard<std::mutex> lk(w->mu);
    const size_t len = w->queue.size() + (size_t)w->inflight;
    if (!best || len < bestLen) {
      best = w;
      bestLen = len;
    }
  }
  {
    std::lock_guard<std::mutex> lk(best->mu);
    best->queue.push_back(req);
  }
  best->cv.notify_one();
  {
    std::unique_lock<std::mutex> lk(req->mu);
    req->cv.wait(lk, [&] { return req->done; });
  }
  if (req->rc) {
    reply_error(fd, req->rc, req->err);
  } else {
    int32_t st = 0;
    uint32_t zero = 0;
    uint64_t n = req->body.size();
    write_full(fd, &st, 4);
    write_full(fd, &zero, 4);
    write_full(fd, &n, 8);
    write_full(fd, req->body.data(), req->body.size());
  }
  ::close(fd);
}

void on_signal(int) {
  g_quit = true;
  if (g_listen >= 0) ::shutdown(g_listen, SHUT_RDWR);
}

int usage() {
  fprintf(stderr, "usage: rmgpu_server --socket PATH [--gpus all|0,1,...] [--max-batch N] [--linger-ms T] [--trace FILE] [--die-with-parent]\n%s\n", rmgpu_version());
  return 2;
}

}  // namespace

int main(int argc, char** argv) {
  std::string gpus = "0";
  for (int a = 1; a < argc; a++) {
    if (!strcmp(argv[a], "--socket") && a + 1 < argc)
      g_socket = argv[++a];
    else if (!strcmp(argv[a], "--gpus") && a + 1 < argc)
      gpus = argv[++a];
    else if (!strcmp(argv[a], "--max-batch") && a + 1 < argc)
      g_maxBatch = std::max(1, atoi(argv[++a]));
    else if (!strcmp(argv[a], "--linger-ms") && a + 1 < argc)
      g_lingerMs = atof(argv[++a]);
    else if (!strcmp(argv[a], "--trace") && a + 1 < argc)
      g_trace = fopen(argv[++a], "a");
    else if (!strcmp(argv[a], "--die-with-parent")) {
      // started by GpuSearchEngine.pm: never outlive the RepeatMasker run that owns this worker
      prctl(PR_SET_PDEATHSIG, SIGTERM);
      if (getppid() == 1) return 0;
    }
    else
      return usage();
  }
  if (g_socket.empty()) return usage();
  if (gpus == "all") {  // every GPU this process can see
    const int n = rmgpu_device_count();
    if (n <= 0) {
      fprintf(stderr, "rmgpu_server: %s\n", rmgpu_last_error());
      return 1;
    }
    gpus.clear();
    for (int i = 0; i < n; i++) gpus += (i ? "," : "") + std::to_string(i);
  }
  for (size_t pos = 0; pos < gpus.size();) {
    size_t comma = gpus.find(',', pos);
    if (comma == std::string::npos) comma = gpus.size();
    auto* w = new Worker();
    w->device = atoi(gpus.substr(pos, comma - pos).c_str());
    g_workers.push_back(w);
    pos = comma + 1;
  }
  signal(SIGPIPE, SIG_IGN);
  signal(SIGINT, on_signal);
  signal(SIGTERM, on_signal);
  ::unlink(g_socket.c_str());
  g_listen = ::socket(AF_UNIX, SOCK_STREAM, 0);
  sockaddr_un sa;
  memset(&sa, 0, sizeof(sa));
  sa.sun_family = AF_UNIX;
  if (g_socket.size() >= sizeof(sa.sun_path)) {
    fprintf(stderr, "rmgpu_server: socket path too long\n");
    return 2;
  }
  strcpy(sa.sun_path, g_socket.c_str());
  if (g_listen < 0 || ::bind(g_listen, (sockaddr*)&sa, sizeof(sa)) != 0 || ::listen(g_listen, 256) != 0) {
    perror("rmgpu_server: bind/listen");
    return 1;
  }
  for (Worker* w : g_workers) w->thread = std::thread(worker_loop, w);
  fprintf(stderr, "rmgpu_server: %s, %zu GPU worker(s), listening on %s\n", rmgpu_version(), g_workers.size(), g_socket.c_str());
  while (!g_quit) {
    int fd = ::accept(g_listen, nullptr, nullptr);
    if (fd < 0) {
      if (g_quit) break;
      continue;
    }
    g_connections++;
    std::thread(handle, fd).detach();  // a connection waits for its own result only
  }
  for (Worker* w : g_workers) w->cv.notify_all();
  for (Worker* w : g_workers) w->thread.join();
  for (int i = 0; i < 2000 && g_connections.load() > 0; i++) std::this_thread::sleep_for(std::chrono::milliseconds(5));
  for (auto* w : g_workers) {
    for (auto& r : w->resident) rmgpu_session_close(r.session);
    fprintf(stderr, "rmgpu_server: GPU %d served %ld searches in %ld passes\n", w->device, w->served, w->passes);
  }
  ::close(g_listen);
  ::unlink(g_socket.c_str());
  return 0;
}
