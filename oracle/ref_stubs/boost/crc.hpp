/* Stand-in for boost::crc_32_type (UDPSinkFEC.cpp:106-109): CRC-32/IEEE. */
#ifndef SDRD_STUB_BOOST_CRC_HPP
#define SDRD_STUB_BOOST_CRC_HPP
#include <cstddef>
#include <cstdint>
extern "C" uint32_t sdro_crc32(const void* data, size_t n);
namespace boost {
class crc_32_type {
public:
    crc_32_type() : m_have(false), m_crc(0) {}
    void process_bytes(const void* p, std::size_t n) { m_crc = sdro_crc32(p, n); m_have = true; }
    uint32_t checksum() const { return m_crc; }
private:
    bool m_have;
    uint32_t m_crc;
};
}
#endif
