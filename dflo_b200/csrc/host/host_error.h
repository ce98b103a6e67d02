// last-error string of the host front end (thread local), see dflo_host_last_error()
#pragma once
#include <string>
namespace dflo
{
   std::string &host_error ();
}
