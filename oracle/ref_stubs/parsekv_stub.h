/* Force-included stand-in for include/parsekv.h (boost::spirit is not in this image).
 * Defines the reference's include guard so its boost version is skipped, and supplies the only
 * symbols the reference uses: parsekv::pairs_type, parsekv::key_value_sequence<> and
 * boost::spirit::qi::parse (DeviceSource.cpp:27-31).  Grammar of parsekv.h:40-43: pairs
 * separated by ',' or '&', key [A-Za-z_][A-Za-z_0-9]*, optional '=' value [A-Za-z_0-9.]+ */
#ifndef INCLUDE_PARSEKV_H_
#define INCLUDE_PARSEKV_H_
#include <cctype>
#include <iostream>
#include <map>
#include <string>
namespace parsekv {
typedef std::map<std::string, std::string> pairs_type;
template <typename Iterator> struct key_value_sequence {};
}
namespace boost { namespace spirit { namespace qi {
template <typename It, typename G>
bool parse(It b, It e, G&, parsekv::pairs_type& m)
{
    while (b != e) {
        std::string k, v;
        if (!(std::isalpha((unsigned char)*b) || *b == '_')) return false;
        while (b != e && (std::isalnum((unsigned char)*b) || *b == '_')) k.push_back(*b++);
        if (b != e && *b == '=') {
            ++b;
            while (b != e && (std::isalnum((unsigned char)*b) || *b == '_' || *b == '.')) v.push_back(*b++);
            if (v.empty()) return false;
        }
        m[k] = v;
        if (b == e) break;
        if (*b == ',' || *b == '&') ++b; else return true; /* qi::parse stops at first mismatch */
    }
    return true;
}
}}}
#endif
