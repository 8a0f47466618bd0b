#pragma once
namespace pi05 {
void set_error(const char* msg);
const char* get_error();
}  // namespace pi05
