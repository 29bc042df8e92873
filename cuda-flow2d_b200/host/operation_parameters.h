// operation_parameters.h -- the reference's by-name parameter bag
// (src/data_types/operation_parameters.{h,cpp}): string -> non-owning void*, first push wins.
#pragma once
#include <string>
#include <unordered_map>

class OperationParameters {
 public:
  bool PushValuePtr(const std::string& key, void* value_ptr) { return map_.emplace(key, value_ptr).second; }
  void* GetValuePtr(const std::string& key) const {
    auto it = map_.find(key);
    return it == map_.end() ? nullptr : it->second;
  }
  void Clear() { map_.clear(); }

 private:
  std::unordered_map<std::string, void*> map_;
};
