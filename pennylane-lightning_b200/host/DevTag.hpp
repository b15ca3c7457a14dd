// Device tag / pool of the B200 backend (mirrors core/utils/cuda_utils/DevTag.hpp and DevicePool.hpp):
// a (device id, stream) pair owned by every state vector, and a thread-safe queue of GPU ids.
#pragma once
#include <algorithm>
#include <cstdint>
#include <deque>
#include <mutex>

#include "../../include/plb200.h"
#include "Error.hpp"

namespace Pennylane::LightningB200 {

template <class IDType = int> class DevTag {
  public:
    DevTag() : device_id_{0}, stream_id_{nullptr} {}
    explicit DevTag(IDType device_id) : device_id_{device_id}, stream_id_{nullptr} {}
    DevTag(IDType device_id, void *stream_id) : device_id_{device_id}, stream_id_{stream_id} {}
    [[nodiscard]] auto getDeviceID() const -> IDType { return device_id_; }
    [[nodiscard]] auto getStreamID() const -> void * { return stream_id_; }
    void refresh() {}
    bool operator==(const DevTag &o) const { return device_id_ == o.device_id_ && stream_id_ == o.stream_id_; }

  private:
    IDType device_id_;
    void *stream_id_;
};

template <class IDType = int> class DevicePool {
  public:
    DevicePool() {
        for (IDType i = 0; i < static_cast<IDType>(getTotalDevices()); i++) available_.push_back(i);
    }
    static std::size_t getTotalDevices() {
        int n = 0;
        if (plb200_device_count(&n) != 0) return 0;
        return static_cast<std::size_t>(n);
    }
    bool isActive(IDType id) {
        std::lock_guard<std::mutex> lk(m_);
        return std::find(available_.begin(), available_.end(), id) == available_.end();
    }
    bool isInactive(IDType id) { return !isActive(id); }
    IDType acquireDevice() {
        std::lock_guard<std::mutex> lk(m_);
        PLB200_ABORT_IF(available_.empty(), "No CUDA device available");
        IDType id = available_.front();
        available_.pop_front();
        return id;
    }
    void releaseDevice(IDType id) {
        std::lock_guard<std::mutex> lk(m_);
        available_.push_back(id);
    }
    void syncDevice(IDType) {}
    void refresh() {
        std::lock_guard<std::mutex> lk(m_);
        available_.clear();
        for (IDType i = 0; i < static_cast<IDType>(getTotalDevices()); i++) available_.push_back(i);
    }

  private:
    std::deque<IDType> available_;
    std::mutex m_;
};

} // namespace Pennylane::LightningB200
