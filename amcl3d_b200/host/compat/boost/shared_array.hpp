// compat stand-in: boost::shared_array<T> with the few members the reference tests use.
#pragma once
#include <memory>
namespace boost
{
template <class T>
class shared_array
{
public:
  shared_array() = default;
  explicit shared_array(T* p) : p_(p, std::default_delete<T[]>()) {}
  T* get() const { return p_.get(); }
  T& operator[](std::ptrdiff_t i) const { return p_.get()[i]; }
  explicit operator bool() const { return static_cast<bool>(p_); }

private:
  std::shared_ptr<T> p_;
};
}  // namespace boost
