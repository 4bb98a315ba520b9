/* Stand-in for boost::crc_32_type (UDPSinkFEC.cpp:106-109): CRC-32/IEEE (reflected 0xEDB88320, initial value and
 * final xor 0xFFFFFFFF), header-only so that reference sources can be compiled with nothing but this header
 * standing in for boost -- against the oracle or against the library under test alike. */
#ifndef SDRD_STUB_BOOST_CRC_HPP
#define SDRD_STUB_BOOST_CRC_HPP
#include <cstddef>
#include <cstdint>
namespace boost {
class crc_32_type {
public:
    crc_32_type() : m_rem(0xFFFFFFFFu) {}
    void process_bytes(const void* p, std::size_t n)
    {
        const unsigned char* b = static_cast<const unsigned char*>(p);
        for (std::size_t i = 0; i < n; i++) {
            m_rem ^= b[i];
            for (int k = 0; k < 8; k++) m_rem = (m_rem & 1u) ? (m_rem >> 1) ^ 0xEDB88320u : m_rem >> 1;
        }
    }
    uint32_t checksum() const { return m_rem ^ 0xFFFFFFFFu; }
private:
    uint32_t m_rem;
};
}
#endif
