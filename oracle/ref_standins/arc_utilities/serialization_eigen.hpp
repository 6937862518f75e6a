/*
 * Stand-in for arc_utilities/serialization_eigen.hpp (+ the serialization.hpp it includes) - TEST INFRASTRUCTURE ONLY
 * (see arc_helpers.hpp next to this file).  The reference's simple_robot_models.hpp needs these names to compile; the byte
 * layouts follow upstream's documented convention as recalled (not verifiable here): fixed-size PODs as
 * their little-endian memory image, vectors as a uint64 element count followed by the elements, fixed-size Eigen types as
 * their coefficients in column-major order, dynamic vectors with a uint64 length in front.
 */
#ifndef SERIALIZATION_EIGEN_STANDIN_HPP
#define SERIALIZATION_EIGEN_STANDIN_HPP

#include <Eigen/Geometry>
#include <cstdint>
#include <cstring>
#include <functional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace arc_utilities
{
template <typename T>
inline uint64_t SerializeFixedSizePOD(const T& item, std::vector<uint8_t>& buffer)
{
    const uint8_t* raw = reinterpret_cast<const uint8_t*>(&item);
    buffer.insert(buffer.end(), raw, raw + sizeof(T));
    return (uint64_t)sizeof(T);
}

template <typename T>
inline std::pair<T, uint64_t> DeserializeFixedSizePOD(const std::vector<uint8_t>& buffer, const uint64_t current)
{
    if (current + sizeof(T) > buffer.size()) throw std::invalid_argument("Not enough room in the provided buffer");
    T item;
    std::memcpy(&item, &buffer[(size_t)current], sizeof(T));
    return std::make_pair(item, (uint64_t)sizeof(T));
}

template <typename T, typename Allocator = std::allocator<T>>
inline uint64_t SerializeVector(const std::vector<T, Allocator>& vec, std::vector<uint8_t>& buffer,
                                const std::function<uint64_t(const T&, std::vector<uint8_t>&)>& item_serializer)
{
    const uint64_t start = (uint64_t)buffer.size();
    SerializeFixedSizePOD<uint64_t>((uint64_t)vec.size(), buffer);
    for (size_t i = 0; i < vec.size(); i++) item_serializer(vec[i], buffer);
    return (uint64_t)buffer.size() - start;
}

template <typename T, typename Allocator = std::allocator<T>>
inline std::pair<std::vector<T, Allocator>, uint64_t> DeserializeVector(const std::vector<uint8_t>& buffer, const uint64_t current,
                                                                        const std::function<std::pair<T, uint64_t>(const std::vector<uint8_t>&, const uint64_t)>& item_deserializer)
{
    uint64_t position = current;
    const std::pair<uint64_t, uint64_t> size = DeserializeFixedSizePOD<uint64_t>(buffer, position);
    position += size.second;
    std::vector<T, Allocator> out;
    out.reserve((size_t)size.first);
    for (uint64_t i = 0; i < size.first; i++)
    {
        const std::pair<T, uint64_t> item = item_deserializer(buffer, position);
        out.push_back(item.first);
        position += item.second;
    }
    return std::make_pair(out, position - current);
}

template <typename CharType>
inline uint64_t SerializeString(const std::basic_string<CharType>& str, std::vector<uint8_t>& buffer)
{
    uint64_t written = SerializeFixedSizePOD<uint64_t>((uint64_t)str.size(), buffer);
    for (size_t i = 0; i < str.size(); i++) written += SerializeFixedSizePOD<CharType>(str[i], buffer);
    return written;
}

template <typename CharType>
inline std::pair<std::basic_string<CharType>, uint64_t> DeserializeString(const std::vector<uint8_t>& buffer, const uint64_t current)
{
    uint64_t position = current;
    const std::pair<uint64_t, uint64_t> size = DeserializeFixedSizePOD<uint64_t>(buffer, position);
    position += size.second;
    std::basic_string<CharType> out;
    for (uint64_t i = 0; i < size.first; i++)
    {
        const std::pair<CharType, uint64_t> c = DeserializeFixedSizePOD<CharType>(buffer, position);
        out.push_back(c.first);
        position += c.second;
    }
    return std::make_pair(out, position - current);
}

/* fixed-size matrices: coefficients, column-major */
template <int R, int C>
inline uint64_t SerializeEigenType(const Eigen::Matrix<double, R, C>& value, std::vector<uint8_t>& buffer)
{
    uint64_t written = 0;
    if (R == Eigen::Dynamic || C == Eigen::Dynamic) written += SerializeFixedSizePOD<uint64_t>((uint64_t)value.size(), buffer);
    for (long j = 0; j < value.cols(); j++)
        for (long i = 0; i < value.rows(); i++) written += SerializeFixedSizePOD<double>(value(i, j), buffer);
    return written;
}

inline uint64_t SerializeEigenType(const Eigen::Isometry3d& value, std::vector<uint8_t>& buffer) { return SerializeEigenType(value.matrix(), buffer); }

template <typename T>
struct EigenTypeDeserializer;

template <int R, int C>
struct EigenTypeDeserializer<Eigen::Matrix<double, R, C>>
{
    static std::pair<Eigen::Matrix<double, R, C>, uint64_t> Run(const std::vector<uint8_t>& buffer, const uint64_t current)
    {
        static_assert(R != Eigen::Dynamic && C != Eigen::Dynamic, "fixed-size only");
        Eigen::Matrix<double, R, C> out;
        uint64_t position = current;
        for (long j = 0; j < C; j++)
            for (long i = 0; i < R; i++)
            {
                const std::pair<double, uint64_t> v = DeserializeFixedSizePOD<double>(buffer, position);
                out(i, j) = v.first;
                position += v.second;
            }
        return std::make_pair(out, position - current);
    }
};

template <>
struct EigenTypeDeserializer<Eigen::VectorXd>
{
    static std::pair<Eigen::VectorXd, uint64_t> Run(const std::vector<uint8_t>& buffer, const uint64_t current)
    {
        uint64_t position = current;
        const std::pair<uint64_t, uint64_t> size = DeserializeFixedSizePOD<uint64_t>(buffer, position);
        position += size.second;
        Eigen::VectorXd out((long)size.first);
        for (uint64_t i = 0; i < size.first; i++)
        {
            const std::pair<double, uint64_t> v = DeserializeFixedSizePOD<double>(buffer, position);
            out((long)i) = v.first;
            position += v.second;
        }
        return std::make_pair(out, position - current);
    }
};

template <>
struct EigenTypeDeserializer<Eigen::Isometry3d>
{
    static std::pair<Eigen::Isometry3d, uint64_t> Run(const std::vector<uint8_t>& buffer, const uint64_t current)
    {
        const std::pair<Eigen::Matrix4d, uint64_t> m = EigenTypeDeserializer<Eigen::Matrix4d>::Run(buffer, current);
        return std::make_pair(Eigen::Isometry3d(m.first), m.second);
    }
};

template <typename T>
inline std::pair<T, uint64_t> DeserializeEigenType(const std::vector<uint8_t>& buffer, const uint64_t current)
{
    return EigenTypeDeserializer<T>::Run(buffer, current);
}
}

#endif
