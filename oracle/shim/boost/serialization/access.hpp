/* shim: friend target for `serialize` members (protocol.hh:66,118,161,185; normaltheta.hh:41). */
#ifndef PF_SHIM_BOOST_SERIALIZATION_ACCESS_HPP
#define PF_SHIM_BOOST_SERIALIZATION_ACCESS_HPP
namespace boost { namespace serialization {
class access {
  public:
    template <typename Archive, typename T>
    static void serialize(Archive& ar, T& t, unsigned version) { t.serialize(ar, version); }
};
} }
#endif
