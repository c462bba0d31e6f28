// Stand-in for Slam_Utility/src/data_struct/circular_buffer.h.
// GUESS G4: a fixed-capacity ring whose PushBack drops the element when full.
#ifndef FD_COMPAT_CIRCULAR_BUFFER_H_
#define FD_COMPAT_CIRCULAR_BUFFER_H_
#include <cstdint>
template <typename T, int32_t kCapacity>
class CircularBuffer {
public:
    void Clear() { head_ = 0; count_ = 0; }
    bool Empty() const { return count_ == 0; }
    bool Full() const { return count_ == kCapacity; }
    int32_t Size() const { return count_; }
    bool PushBack(const T &v) {
        if (count_ == kCapacity) return false;
        slots_[(head_ + count_) % kCapacity] = v;
        ++count_;
        return true;
    }
    T &Front() { return slots_[head_]; }
    const T &Front() const { return slots_[head_]; }
    void PopFront() {
        if (count_ == 0) return;
        head_ = (head_ + 1) % kCapacity;
        --count_;
    }
private:
    T slots_[kCapacity];
    int32_t head_ = 0;
    int32_t count_ = 0;
};
#endif  // FD_COMPAT_CIRCULAR_BUFFER_H_
