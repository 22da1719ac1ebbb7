// kdpt_bulk.h — std::vector without value-initialisation for the per-entry arrays (hundreds of MB to ~1 GB on the large
// configs). Pure host C++.
#pragma once
#include <memory>
#include <new>
#include <utility>
#include <vector>

namespace kdpt {

// Every element of such an array is written by a parallel fill (or a device-to-host copy) right after the resize; a
// zero-filling resize would first fault the whole array in on one thread.
template <class T> struct DefaultInitAllocator : std::allocator<T> {
    template <class U> struct rebind { typedef DefaultInitAllocator<U> other; };
    using std::allocator<T>::allocator;
    template <class U> void construct(U* p) noexcept { ::new ((void*)p) U; }
    template <class U, class... A> void construct(U* p, A&&... a) { ::new ((void*)p) U(std::forward<A>(a)...); }
};
template <class T> using BulkVector = std::vector<T, DefaultInitAllocator<T>>;

} // namespace kdpt
