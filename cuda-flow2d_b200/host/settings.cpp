#include "settings.h"

#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>
#include <vector>

namespace OpticFlow {
namespace {

struct Element {
  std::string path;  // e.g. "OpticalFlow/Parameters/Solver/Model"
  std::map<std::string, std::string> attr;
};

// Minimal scanner: start tags with attributes, end tags, comments, declarations.  No entities, no
// CDATA, no text content -- settings.xml has none.
bool scan(const std::string& s, std::vector<Element>* out, std::string* err) {
  std::vector<std::string> stack;
  size_t i = 0;
  while ((i = s.find('<', i)) != std::string::npos) {
    if (s.compare(i, 4, "<!--") == 0) {
      size_t e = s.find("-->", i);
      if (e == std::string::npos) { *err = "unterminated comment"; return false; }
      i = e + 3;
      continue;
    }
    if (s.compare(i, 2, "<?") == 0 || s.compare(i, 2, "<!") == 0) {
      size_t e = s.find('>', i);
      if (e == std::string::npos) { *err = "unterminated declaration"; return false; }
      i = e + 1;
      continue;
    }
    size_t e = i + 1;
    char quote = 0;
    for (; e < s.size(); ++e) {  // find the closing '>' outside quotes
      if (quote) { if (s[e] == quote) quote = 0; }
      else if (s[e] == '"' || s[e] == '\'') quote = s[e];
      else if (s[e] == '>') break;
    }
    if (e >= s.size()) { *err = "unterminated tag"; return false; }
    std::string tag = s.substr(i + 1, e - i - 1);
    i = e + 1;
    if (!tag.empty() && tag[0] == '/') {
      if (stack.empty()) { *err = "unbalanced end tag"; return false; }
      stack.pop_back();
      continue;
    }
    bool self_closing = !tag.empty() && tag.back() == '/';
    if (self_closing) tag.pop_back();
    size_t p = 0;
    while (p < tag.size() && !isspace((unsigned char)tag[p])) ++p;
    Element el;
    std::string name = tag.substr(0, p);
    for (const std::string& a : stack) el.path += a + "/";
    el.path += name;
    while (p < tag.size()) {  // attributes: name [ws] = [ws] "value"
      while (p < tag.size() && isspace((unsigned char)tag[p])) ++p;
      size_t n0 = p;
      while (p < tag.size() && !isspace((unsigned char)tag[p]) && tag[p] != '=') ++p;
      std::string an = tag.substr(n0, p - n0);
      while (p < tag.size() && isspace((unsigned char)tag[p])) ++p;
      if (an.empty() || p >= tag.size() || tag[p] != '=') break;
      ++p;
      while (p < tag.size() && isspace((unsigned char)tag[p])) ++p;
      if (p >= tag.size() || (tag[p] != '"' && tag[p] != '\'')) { *err = "attribute value must be quoted"; return false; }
      char q = tag[p++];
      size_t v0 = p;
      while (p < tag.size() && tag[p] != q) ++p;
      el.attr[an] = tag.substr(v0, p - v0);
      ++p;
    }
    out->push_back(el);
    if (!self_closing) stack.push_back(name);
  }
  return true;
}

}  // namespace

int Settings::LoadSettings(const std::string& fileName) {
  std::ifstream f(fileName.c_str());
  if (!f) {
    error = "Cannot read settings file: " + fileName;
    return -1;
  }
  std::stringstream buffer;
  buffer << f.rdbuf();
  std::vector<Element> els;
  if (!scan(buffer.str(), &els, &error) || els.empty()) {
    if (error.empty()) error = "Problem with parsing settings file: " + fileName;
    return -1;
  }
  const std::string root = els[0].path;
  bool ok = true;
  auto attr = [&](const std::string& path, const std::string& name, bool required, std::string* out) {
    for (const Element& e : els)
      if (e.path == root + "/" + path) {
        auto it = e.attr.find(name);
        if (it != e.attr.end()) { *out = it->second; return true; }
      }
    if (required) { ok = false; error = "missing " + path + "@" + name; }
    return false;
  };
  std::string v;
  attr("Input/Path", "inputPath", true, &inputPath);
  attr("Output/Path", "outputPath", true, &outputPath);
  attr("Input/Mode/Files", "file1", true, &fileName1);
  attr("Input/Mode/Files", "file2", true, &fileName2);
  attr("Input/Mode", "imageType", false, &imageType);
  if (attr("Parameters/Method", "key", false, &v)) press_key = std::atoi(v.c_str()) != 0;
  if (attr("Input/Mode", "Nx", true, &v)) width = std::atoi(v.c_str());
  if (attr("Input/Mode", "Ny", true, &v)) height = std::atoi(v.c_str());
  if (attr("Parameters/Solver/Model", "sigma", true, &v)) sigma = std::strtof(v.c_str(), nullptr);
  if (attr("Parameters/Solver/Iterations", "inner", true, &v)) iterInner = std::atoi(v.c_str());
  if (attr("Parameters/Solver/Iterations", "outer", true, &v)) iterOuter = std::atoi(v.c_str());
  if (attr("Parameters/Solver/Warping", "levels", true, &v)) levels = std::atoi(v.c_str());
  if (attr("Parameters/Solver/Warping", "scaling", true, &v)) warpScale = std::strtof(v.c_str(), nullptr);
  if (attr("Parameters/Solver/Warping", "medianRadius", true, &v)) medianRadius = std::atoi(v.c_str());
  if (attr("Parameters/Solver/Model", "alpha", true, &v)) alpha = std::strtof(v.c_str(), nullptr);
  if (attr("Parameters/Solver/Model", "e_smooth", true, &v)) e_smooth = std::strtof(v.c_str(), nullptr);
  if (attr("Parameters/Solver/Model", "e_data", true, &v)) e_data = std::strtof(v.c_str(), nullptr);
  attr("Parameters/Solver/Model", "constancy", false, &constancy);
  return ok ? 0 : -1;
}

}  // namespace OpticFlow
