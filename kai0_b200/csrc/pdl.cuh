// Device side of programmatic dependent launch (see launch.h for the contract).
#pragma once

namespace pi05 {

// Blocks until every kernel this one depends on has completed and its global writes are visible.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Lets the next kernel of the stream be staged (it still waits for THIS kernel's completion in its own pdl_wait()).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() {
  pdl_wait();
  pdl_launch_dependents();
}

}  // namespace pi05
