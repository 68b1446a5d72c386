// compat stand-in: boost::shared_ptr mapped onto std::shared_ptr (see compat/README.md).
#pragma once
#include <memory>
namespace boost
{
template <class T>
using shared_ptr = std::shared_ptr<T>;
template <class T, class... Args>
inline std::shared_ptr<T> make_shared(Args&&... args)
{
  return std::make_shared<T>(std::forward<Args>(args)...);
}
}  // namespace boost
