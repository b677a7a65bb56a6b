// Implementation of the small tinyxml2 stand-in declared in tinyxml2.h (oracle/_ref build only).
#include "tinyxml2.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace tinyxml2 {

// ---------------------------------------------------------------- XMLNode

XMLNode::~XMLNode() { DeleteChildren(); }

void XMLNode::Unlink(XMLNode* c) {
  if (c->prev_) c->prev_->next_ = c->next_; else first_ = c->next_;
  if (c->next_) c->next_->prev_ = c->prev_; else last_ = c->prev_;
  c->prev_ = c->next_ = nullptr;
  c->parent_ = nullptr;
}

void XMLNode::DeleteChildren() {
  while (first_) {
    XMLNode* c = first_;
    Unlink(c);
    delete c;
  }
}

void XMLNode::DeleteChild(XMLNode* node) {
  if (!node || node->parent_ != this) return;
  Unlink(node);
  delete node;
}

XMLNode* XMLNode::InsertEndChild(XMLNode* add) {
  if (!add) return nullptr;
  if (add->parent_) add->parent_->Unlink(add);
  add->parent_ = this;
  add->prev_ = last_;
  add->next_ = nullptr;
  if (last_) last_->next_ = add; else first_ = add;
  last_ = add;
  return add;
}

XMLNode* XMLNode::InsertFirstChild(XMLNode* add) {
  if (!add) return nullptr;
  if (add->parent_) add->parent_->Unlink(add);
  add->parent_ = this;
  add->prev_ = nullptr;
  add->next_ = first_;
  if (first_) first_->prev_ = add; else last_ = add;
  first_ = add;
  return add;
}

XMLNode* XMLNode::InsertAfterChild(XMLNode* after, XMLNode* add) {
  if (!add || !after || after->parent_ != this) return nullptr;
  if (after == add) return add;
  if (!after->next_) return InsertEndChild(add);
  if (add->parent_) add->parent_->Unlink(add);
  add->parent_ = this;
  add->prev_ = after;
  add->next_ = after->next_;
  after->next_->prev_ = add;
  after->next_ = add;
  return add;
}

static bool NameIs(const XMLNode* n, const char* name) {
  return !name || std::strcmp(n->Value(), name) == 0;
}

XMLElement* XMLNode::FirstChildElement(const char* name) {
  for (XMLNode* n = first_; n; n = n->next_) {
    XMLElement* e = n->ToElement();
    if (e && NameIs(e, name)) return e;
  }
  return nullptr;
}
const XMLElement* XMLNode::FirstChildElement(const char* name) const {
  return const_cast<XMLNode*>(this)->FirstChildElement(name);
}
XMLElement* XMLNode::NextSiblingElement(const char* name) {
  for (XMLNode* n = next_; n; n = n->next_) {
    XMLElement* e = n->ToElement();
    if (e && NameIs(e, name)) return e;
  }
  return nullptr;
}
const XMLElement* XMLNode::NextSiblingElement(const char* name) const {
  return const_cast<XMLNode*>(this)->NextSiblingElement(name);
}

XMLNode* XMLNode::DeepClone(XMLDocument* target) const {
  XMLNode* c = ShallowClone(target);
  if (!c) return nullptr;
  for (const XMLNode* n = first_; n; n = n->next_) {
    XMLNode* cc = n->DeepClone(target);
    if (cc) c->InsertEndChild(cc);
  }
  return c;
}

// ---------------------------------------------------------------- XMLComment / XMLElement

XMLNode* XMLComment::ShallowClone(XMLDocument* doc) const {
  XMLComment* c = new XMLComment(doc);
  c->value_ = value_;
  c->line_ = line_;
  return c;
}

XMLElement::~XMLElement() {
  while (attr_) {
    XMLAttribute* a = attr_;
    attr_ = a->next_;
    delete a;
  }
}

XMLNode* XMLElement::ShallowClone(XMLDocument* doc) const {
  XMLElement* e = new XMLElement(doc);
  e->value_ = value_;
  e->line_ = line_;
  for (const XMLAttribute* a = attr_; a; a = a->next_) e->SetAttribute(a->Name(), a->Value());
  return e;
}

const XMLAttribute* XMLElement::FindAttribute(const char* name) const {
  for (const XMLAttribute* a = attr_; a; a = a->next_) {
    if (a->name_ == name) return a;
  }
  return nullptr;
}

const char* XMLElement::Attribute(const char* name, const char* value) const {
  const XMLAttribute* a = FindAttribute(name);
  if (!a) return nullptr;
  if (!value || a->value_ == value) return a->Value();
  return nullptr;
}

void XMLElement::SetAttribute(const char* name, const char* value) {
  XMLAttribute* last = nullptr;
  for (XMLAttribute* a = attr_; a; a = a->next_) {
    if (a->name_ == name) { a->value_ = value; return; }
    last = a;
  }
  XMLAttribute* a = new XMLAttribute;
  a->name_ = name;
  a->value_ = value;
  if (last) last->next_ = a; else attr_ = a;
}

void XMLElement::SetAttribute(const char* name, int value) {
  char buf[64];
  std::snprintf(buf, sizeof(buf), "%d", value);
  SetAttribute(name, buf);
}

void XMLElement::SetAttribute(const char* name, double value) {
  char buf[64];
  std::snprintf(buf, sizeof(buf), "%.17g", value);
  SetAttribute(name, buf);
}

void XMLElement::DeleteAttribute(const char* name) {
  XMLAttribute* prev = nullptr;
  for (XMLAttribute* a = attr_; a; prev = a, a = a->next_) {
    if (a->name_ == name) {
      if (prev) prev->next_ = a->next_; else attr_ = a->next_;
      delete a;
      return;
    }
  }
}

// ---------------------------------------------------------------- XMLDocument

XMLElement* XMLDocument::NewElement(const char* name) {
  XMLElement* e = new XMLElement(this);
  e->value_ = name;
  return e;
}

XMLComment* XMLDocument::NewComment(const char* text) {
  XMLComment* c = new XMLComment(this);
  c->value_ = text;
  return c;
}

void XMLDocument::Fail(int line, const std::string& msg) {
  err_ = XML_ERROR_PARSING;
  errline_ = line;
  char buf[64];
  std::snprintf(buf, sizeof(buf), " line=%d", line);
  errstr_ = "Error=XML_ERROR_PARSING ErrorID=1 (0x1)" + std::string(buf) + ": " + msg;
}

namespace {

struct Cursor {
  const char* p;
  const char* end;
  int line;
  bool Eof() const { return p >= end; }
  char Peek() const { return p < end ? *p : '\0'; }
  void Advance(size_t n = 1) {
    while (n-- && p < end) {
      if (*p == '\n') ++line;
      ++p;
    }
  }
  bool StartsWith(const char* s) const {
    size_t n = std::strlen(s);
    return (size_t)(end - p) >= n && std::strncmp(p, s, n) == 0;
  }
  void SkipSpace() {
    while (p < end && (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\n')) Advance();
  }
  // advance past the first occurrence of s; false if not found
  bool SkipPast(const char* s) {
    while (!Eof()) {
      if (StartsWith(s)) { Advance(std::strlen(s)); return true; }
      Advance();
    }
    return false;
  }
};

bool IsNameChar(char c) {
  return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9') ||
         c == '_' || c == ':' || c == '-' || c == '.' || (unsigned char)c >= 128;
}

void AppendUtf8(std::string& out, unsigned long cp) {
  if (cp < 0x80) out += (char)cp;
  else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
  else if (cp < 0x10000) {
    out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F));
    out += (char)(0x80 | (cp & 0x3F));
  } else {
    out += (char)(0xF0 | (cp >> 18)); out += (char)(0x80 | ((cp >> 12) & 0x3F));
    out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F));
  }
}

std::string DecodeEntities(const char* s, size_t n) {
  std::string out;
  out.reserve(n);
  for (size_t i = 0; i < n; ++i) {
    if (s[i] != '&') { out += s[i]; continue; }
    const char* semi = (const char*)std::memchr(s + i, ';', n - i);
    if (!semi) { out += s[i]; continue; }
    std::string ent(s + i + 1, semi - (s + i + 1));
    if (ent == "lt") out += '<';
    else if (ent == "gt") out += '>';
    else if (ent == "amp") out += '&';
    else if (ent == "quot") out += '"';
    else if (ent == "apos") out += '\'';
    else if (!ent.empty() && ent[0] == '#') {
      unsigned long cp = (ent.size() > 1 && (ent[1] == 'x' || ent[1] == 'X'))
                             ? std::strtoul(ent.c_str() + 2, nullptr, 16)
                             : std::strtoul(ent.c_str() + 1, nullptr, 10);
      AppendUtf8(out, cp);
    } else { out += s[i]; continue; }
    i = semi - s;
  }
  return out;
}

}  // namespace

XMLError XMLDocument::Parse(const char* xml, size_t nbytes) {
  Clear();
  if (!xml) { Fail(0, "empty document"); return err_; }
  if (nbytes == static_cast<size_t>(-1)) nbytes = std::strlen(xml);
  Cursor c{xml, xml + nbytes, 1};
  // a trailing NUL inside the buffer ends the document
  const char* nul = (const char*)std::memchr(xml, 0, nbytes);
  if (nul) c.end = nul;
  if (c.StartsWith("\xEF\xBB\xBF")) c.Advance(3);

  XMLNode* cur = this;
  while (true) {
    // text between tags is dropped
    while (!c.Eof() && c.Peek() != '<') c.Advance();
    if (c.Eof()) break;
    int line = c.line;
    if (c.StartsWith("<!--")) {
      c.Advance(4);
      const char* s = c.p;
      if (!c.SkipPast("-->")) { Fail(line, "unterminated comment"); return err_; }
      XMLComment* cm = new XMLComment(this);
      cm->value_.assign(s, c.p - 3 - s);
      cm->line_ = line;
      cur->InsertEndChild(cm);
    } else if (c.StartsWith("<![CDATA[")) {
      if (!c.SkipPast("]]>")) { Fail(line, "unterminated CDATA"); return err_; }
    } else if (c.StartsWith("<?")) {
      if (!c.SkipPast("?>")) { Fail(line, "unterminated declaration"); return err_; }
    } else if (c.StartsWith("<!")) {
      if (!c.SkipPast(">")) { Fail(line, "unterminated <! block"); return err_; }
    } else if (c.StartsWith("</")) {
      c.Advance(2);
      const char* s = c.p;
      while (IsNameChar(c.Peek())) c.Advance();
      std::string name(s, c.p - s);
      c.SkipSpace();
      if (c.Peek() != '>') { Fail(line, "malformed closing tag"); return err_; }
      c.Advance();
      if (cur == this || name != cur->Value()) {
        Fail(line, "mismatched closing tag </" + name + ">");
        return err_;
      }
      cur = cur->Parent();
    } else {
      c.Advance();  // '<'
      const char* s = c.p;
      while (IsNameChar(c.Peek())) c.Advance();
      if (c.p == s) { Fail(line, "element name expected"); return err_; }
      XMLElement* e = new XMLElement(this);
      e->value_.assign(s, c.p - s);
      e->line_ = line;
      cur->InsertEndChild(e);
      bool closed = false, done = false;
      while (!done) {
        c.SkipSpace();
        if (c.Eof()) { Fail(line, "unterminated element"); return err_; }
        if (c.StartsWith("/>")) { c.Advance(2); closed = true; done = true; }
        else if (c.Peek() == '>') { c.Advance(); done = true; }
        else {
          const char* as = c.p;
          while (IsNameChar(c.Peek())) c.Advance();
          if (c.p == as) { Fail(c.line, "attribute name expected"); return err_; }
          std::string aname(as, c.p - as);
          c.SkipSpace();
          if (c.Peek() != '=') { Fail(c.line, "'=' expected after attribute " + aname); return err_; }
          c.Advance();
          c.SkipSpace();
          char q = c.Peek();
          if (q != '"' && q != '\'') { Fail(c.line, "quoted attribute value expected"); return err_; }
          c.Advance();
          const char* vs = c.p;
          while (!c.Eof() && c.Peek() != q) c.Advance();
          if (c.Eof()) { Fail(c.line, "unterminated attribute value"); return err_; }
          std::string val = DecodeEntities(vs, c.p - vs);
          c.Advance();
          if (e->FindAttribute(aname.c_str())) { Fail(c.line, "duplicate attribute " + aname); return err_; }
          e->SetAttribute(aname.c_str(), val.c_str());
        }
      }
      if (!closed) cur = e;
    }
  }
  if (cur != this) { Fail(c.line, std::string("unclosed element <") + cur->Value() + ">"); return err_; }
  if (!RootElement()) { Fail(c.line, "no root element"); return err_; }
  return err_;
}

// ---------------------------------------------------------------- XMLPrinter

static std::string Escape(const char* s) {
  std::string out;
  for (; *s; ++s) {
    switch (*s) {
      case '<': out += "&lt;"; break;
      case '>': out += "&gt;"; break;
      case '&': out += "&amp;"; break;
      case '"': out += "&quot;"; break;
      default: out += *s;
    }
  }
  return out;
}

void XMLPrinter::PrintNode(const XMLNode* node, int depth) {
  if (const XMLElement* e = node->ToElement()) {
    if (!compact_) PrintSpace(depth);
    Write("<");
    Write(e->Name());
    for (const XMLAttribute* a = e->FirstAttribute(); a; a = a->Next()) {
      Write(" ");
      Write(a->Name());
      Write("=\"");
      Write(Escape(a->Value()).c_str());
      Write("\"");
    }
    if (e->NoChildren()) {
      Write("/>");
      if (!compact_) Write("\n");
    } else {
      Write(">");
      if (!compact_) Write("\n");
      for (const XMLNode* c = e->FirstChild(); c; c = c->NextSibling()) PrintNode(c, depth + 1);
      if (!compact_) PrintSpace(depth);
      Write("</");
      Write(e->Name());
      Write(">");
      if (!compact_) Write("\n");
    }
  } else if (node->ToComment()) {
    if (!compact_) PrintSpace(depth);
    Write("<!--");
    Write(node->Value());
    Write("-->");
    if (!compact_) Write("\n");
  }
}

void XMLDocument::Print(XMLPrinter* printer) const {
  XMLPrinter local;
  XMLPrinter* p = printer ? printer : &local;
  for (const XMLNode* c = FirstChild(); c; c = c->NextSibling()) p->PrintNode(c, 0);
  if (!printer) std::fputs(local.CStr(), stdout);
}

}  // namespace tinyxml2
