// Launch timeline of the multi-stream drivers (developer tool; off unless gpb_trace_begin() was called).
//
// A TraceSpan brackets ONE launch (or one exchange) with two CUDA events recorded on the launch's stream: the first event
// completes when the stream reaches the launch (everything before it on that stream is done), the second when the launch
// has finished.  The span therefore contains the time the launch waited for SM resources - exactly what the look-ahead
// schedules of run_potrf / run_potrf_dist need to be judged by (nsys is not available on the pool).  Events cannot be
// timed inside a stream capture: trace gpb_plan_eval, not the graph replay of gpb_plan_eval_host.
#pragma once
#include <cuda_runtime.h>

namespace gpb {

bool trace_active();
// id of the span or -1; tag = short static string, a / b = free integers (block column, rows, ...)
int trace_open(const char* tag, cudaStream_t s, int a, int b);
void trace_close(int id, cudaStream_t s);

struct TraceSpan {
  int id;
  cudaStream_t s;
  TraceSpan(const char* tag, cudaStream_t s_, int a = 0, int b = 0) : id(-1), s(s_) {
    if (trace_active()) id = trace_open(tag, s_, a, b);
  }
  ~TraceSpan() { if (id >= 0) trace_close(id, s); }
};

}  // namespace gpb
