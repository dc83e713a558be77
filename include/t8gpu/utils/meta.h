/// @file meta.h  --  compile-time helpers with the names of t8gpu/utils/meta.h:25-120 (fold-expression based).
#ifndef T8GPU_B200_UTILS_META_H
#define T8GPU_B200_UTILS_META_H

#include <cstddef>
#include <type_traits>
#include <utility>

namespace t8gpu::meta {

  /// true when all types are the same up to cv-qualification.
  template<typename... Ts>
  struct all_same : std::false_type {};
  template<typename T, typename... Ts>
  struct all_same<T, Ts...> : std::bool_constant<(std::is_same_v<std::remove_cv_t<T>, std::remove_cv_t<Ts>> && ...)> {};
  template<typename... Ts>
  inline constexpr bool all_same_v = all_same<Ts...>::value;

  /// true when static_cast<U>(T) is well formed.
  template<typename T, typename U, typename = void>
  struct is_explicitly_convertible_to : std::false_type {};
  template<typename T, typename U>
  struct is_explicitly_convertible_to<T, U, std::void_t<decltype(static_cast<U>(std::declval<T>()))>> : std::true_type {};
  template<typename T, typename U>
  inline constexpr bool is_explicitly_convertible_to_v = is_explicitly_convertible_to<T, U>::value;

  namespace detail {
    template<int... args>
    constexpr int pack_at(int index) {
      constexpr int vals[] = {args...};
      return vals[index];
    }
    template<int... args>
    constexpr int pack_mul(int lo, int hi) {  // product of entries lo <= i < hi
      constexpr int vals[] = {args...};
      int           r      = 1;
      for (int i = lo; i < hi; i++) r *= vals[i];
      return r;
    }
  }  // namespace detail

  /// index-th value of an integer pack.
  template<int index, int... args>
  struct argpack_at : std::integral_constant<int, detail::pack_at<args...>(index)> {};
  template<int index, int... args>
  inline constexpr int argpack_at_v = argpack_at<index, args...>::value;

  /// product of the values with position >= index.
  template<int index, int... args>
  struct argpack_mul_from : std::integral_constant<int, detail::pack_mul<args...>(index, sizeof...(args))> {};
  template<int index, int... args>
  inline constexpr int argpack_mul_from_v = argpack_mul_from<index, args...>::value;

  /// product of the values with position < index (the column-major stride of dimension `index`).
  template<int index, int... args>
  struct argpack_mul_to : std::integral_constant<int, detail::pack_mul<args...>(0, index)> {};
  template<int index, int... args>
  inline constexpr int argpack_mul_to_v = argpack_mul_to<index, args...>::value;

  /// floor(log2(x)).
  template<size_t x>
  struct log2 : std::integral_constant<size_t, 1 + log2<x / 2>::value> {};
  template<>
  struct log2<1> : std::integral_constant<size_t, 0> {};
  template<size_t x>
  inline constexpr size_t log2_v = log2<x>::value;

}  // namespace t8gpu::meta

#endif  // T8GPU_B200_UTILS_META_H
