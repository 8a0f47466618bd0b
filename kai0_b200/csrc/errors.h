#pragma once
namespace pi05 {
void set_error(const char* msg);
const char* get_error();
void count_launch();          // one per kernel launch of this library (bench.py's gpu_launches)
unsigned long long launch_count();
}  // namespace pi05
